// K2/K3/K4: the theta-rule diffusion step as ONE persistent cooperative kernel per time step.
//
// Replaces, from the reference's BaseModel.step (src/beat/base_model.py:208-245):
//   _update_rhs  (FFCx cell loop + assemble_vector, :196-206)  -> b = B v_ + dt * sum_k a_k(t) s_k   (K2)
//   _update_matrices (assemble_matrix when dt changes, :188-194,:225-230) -> A,B from Mass,K          (K3)
//   KSP solve (PETSc cg + hypre / LU, :236)                      -> Jacobi-preconditioned CG           (K4)
// with A = C_m*Mass + dt*theta*K, B = C_m*Mass - dt*(1-theta)*K (form: monodomain_model.py:83-96).
//
// Storage: SELL-32 (sliced ELLPACK, slice = 32 rows = one warp, column-major inside a slice) built once
// from the host CSR; A and B share the column indices.  Lane r of a warp owns row 32*s+r, so every
// value/index load is one fully coalesced 256 B / 128 B transaction; padding has value 0 and the row's
// own column.  Row ownership is static: thread `tid` of the grid owns rows tid, tid+T, tid+2T, ... in
// every phase of the solve, so everything a phase produces for its own row stays in shared memory (or
// in thread-private global arrays for large meshes) and only the ONE vector that neighbour rows gather
// (the search direction) is exchanged between CTAs.
//
// The whole solve (RHS, stimulus, initial residual, every CG iteration, convergence test) runs inside
// one cooperative launch of one 512-thread CTA per SM, and it contains NO memory fence and NO atomic:
// measured on B200 (tests/cuda/sync_bench.cu) a gpu-scope fence costs about 1.2 us, so a classic
// fence + flag barrier is 2.5 us (cooperative-groups grid.sync) to 6 us.  Instead every value that
// crosses CTAs travels WITH its generation number in one aligned 16-byte store {double, uint64}
// (the idea of NCCL's LL protocol):
//   * the exchanged vector is an array of such pairs; a gather spins on the tag of the element it
//     needs - dataflow synchronisation at element granularity, no barrier before the SpMV;
//   * the dot products: every CTA publishes tagged partial sums, CTA 0 polls them (one record per
//     thread), adds them in a fixed order and publishes tagged totals that everybody polls:
//     1.9 us per reduction, bit-reproducible, and all CTAs see the same totals so they take the same
//     branch on the convergence test.
// Two buffers alternate for the exchanged vector; the reduction between two writes of the same buffer
// is an all-to-all dependency, which is what makes the reuse safe (DESIGN.md "K4").
//
// Two Krylov drivers:
//   pde_cg_kernel      KSPCG as PETSc runs it: two reductions per iteration.
//   pde_pipecg_kernel  pipelined CG (Ghysels & Vanroose; PETSc's KSPPIPECG): mathematically the same
//                      iterates, ONE reduction per iteration.
// RESIDENT=true keeps the CG vectors of a CTA's rows in shared memory (up to 3072 rows per SM = 454k
// rows per GPU); RESIDENT=false streams them from global memory (HBM-bound for large meshes).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "mono_ctx.h"

namespace {

constexpr int kSlice = 32;
constexpr int kPdeThreads = 512;
constexpr int kWarpsPerBlock = kPdeThreads / 32;
constexpr int kChunk = 16;            // SELL entries of a row fetched per unrolled batch
constexpr int kMaxBlocks = 160;       // >= SM count: size of the reduction scratch in shared memory
constexpr int kMaxPat = 32;           // stencil dictionary: patterns (of at most kChunk entries) held in shared memory
constexpr int kDictSmem = kMaxPat * kChunk * (8 + 8 + 4) + kMaxPat * 4;  // A values, B values, offsets, widths
constexpr int kRingRows = 2;                           // rows per thread and tile
constexpr int kRingTile = kRingRows * kPdeThreads;     // rows of one CTA trip of the ring kernel
constexpr int kRingChunk = kRingTile;                  // elements per bulk copy into a ring (rings advance in whole chunks)
constexpr int kRingChunkLog2 = 10;
static_assert((1 << kRingChunkLog2) == kRingChunk, "chunk size");
constexpr int kRingMinDepth = 1, kRingMaxDepth = 6;    // tiles whose ring data is requested ahead of the one being computed
constexpr int kRingStages = kRingMaxDepth + 1;         // mbarriers (one per tile in flight)
constexpr int kMaxClusters = 4;
constexpr int kRingTableSmem = kMaxPat * kChunk * (8 + 8 + 8) + kMaxPat * (4 + 8);  // A, B, {offset, cluster ring} pairs, widths, 1/diagonal
constexpr int kRingStageBytes = kRingTile + kRingTile / kSlice;  // per tile in flight: its pattern bytes and the stimulus flags of its slices
typedef unsigned long long u64;

// where one owned boundary value goes on ONE neighbour rank (peer-mapped addresses of its ghost slot)
struct SendEnt {
  SyncRec* t[kMaxTb];
  SyncRec* xg;
  int32_t next;  // next entry of the same row (a dof can be a ghost on several ranks), -1 ends the list
  int32_t pad;
};

struct PdeArgs {
  int64_t n_owned, n_local, n_slices;
  const int64_t* slice_ptr;
  const int32_t* cols;
  const double* A;
  const double* B;
  const double* dinv;
  const double* v_prev;
  double* x;                        // solution (owned rows written here at the end)
  double* work[10];                 // streaming mode: thread-private vectors (n_owned each)
  SyncRec* tb[kMaxTb];              // exchanged vectors, tagged {value, generation}: 2 buffers of n_local (Jacobi), or
                                    // 3 sets of cheb_k buffers (Chebyshev: one buffer per polynomial step)
  int cheb_k;                       // Chebyshev steps of the preconditioner (1 = plain Jacobi)
  double cheb_inv_theta;            // y_1 = g / theta
  double cheb_c1[kMaxCheb], cheb_c2[kMaxCheb];  // d_j = c1_j d_{j-1} + c2_j (g - D^-1 A y_j), y_{j+1} = y_j + d_j
  const double* stim_vec;           // dense sum_k a_k(t) s_k over owned rows (valid when has_stim)
  int has_stim;
  int staged;                       // streaming mode with every slice at most kChunk wide: TMA-staged SpMV
  int rows_per_thread;              // ceil(n_slices / warps of the worker CTAs)
  int n_workers;                    // CTAs 0 .. n_workers-1 own rows; CTA n_workers only reduces
  double dt;
  double rtol, atol;
  int max_it, norm_type, x0_mode;
  SyncRec* recs;                    // partial sums [2 parities][4 slots][n_workers], then totals [2][4]
  u64* gen_state;                   // [0] next unused generation number (device-resident: continues across launches,
                                    //     so every rank of a multi-GPU run counts the same generations)
  u64 spin_ns;                      // a poll gives up after this long (a peer died): the launch reports an error
  // ---- multi-GPU (one rank per GPU; peers' buffers are mapped with CUDA IPC, stores travel over NVLink) ----
  int nranks, rank;
  SyncRec* xg;                      // landing zone of the neighbours' final x values (n_ghost, tagged)
  SyncRec* peer_xrecs[kMaxRanks];   // every rank's cross-rank reduction records [2 parities][4 slots][nranks]
  const int32_t* send_of_row;       // owned row -> first entry of its send list (-1: interior row)
  const SendEnt* send_ents;
  const uint8_t* slice_send;        // per SELL slice: 1 when some row of it has a send list (streaming kernel: skips the lookups)
  KspResult* res;
  u64* timeline;                    // optional (measurement): globaltimer stamps of CTA 0 at phase boundaries
  // ---- stencil dictionary (DICT kernels only) ----
  const uint8_t* pat;               // n_owned: pattern of a row, 255 = take the row from the SELL arrays
  const double* dict_A;             // [n_pat][kChunk]
  const double* dict_B;
  const int32_t* dict_off;          // [n_pat][kChunk] column - row
  const int32_t* dict_w;            // [n_pat]
  int n_pat;
  // ---- shared-memory ring of the dictionary SpMV (MODE 2) and the side table of the rows outside the dictionary ----
  const int32_t* dict_cl;           // [n_pat][kChunk] cluster of an entry
  int ring_ncl, ring_cap_log2, ring_cl0, ring_depth;  // clusters, ring capacity, cluster of offset 0, tiles requested ahead
  int64_t ring_lo[kMaxClusters], ring_hi[kMaxClusters];
  const int32_t* nd_rows;           // [n_nd]
  const int32_t* nd_w;
  const int32_t* nd_cols;           // [n_nd][kChunk]
  const double* nd_A;
  const double* nd_B;
  const int64_t* nd_cta_ptr;        // [n_workers + 1]
  const double* dict_dinv;          // [n_pat] Jacobi diagonal of a stencil (= dinv of any row that has it)
  const uint8_t* slice_stim;        // per SELL slice: 1 when stim_vec is non-zero on some row of it
  int dbg;                          // measurement only (MONO_RING_DBG): 1 = no ring requests, 2 = no ring compute, 4 = no q store
};

__device__ __forceinline__ void stamp(const PdeArgs& a, int& n) {
  if (a.timeline != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && n < 63) {
    u64 t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.timeline[1 + n++] = t;
    a.timeline[0] = (u64)n;
  }
}

struct Scratch {                    // static shared memory of the persistent kernels
  double warp_part[4 * kWarpsPerBlock];
  double totals[4];
  double red[4 * kMaxBlocks];
  int fail;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One 16-byte store / load: value and tag always travel together.  The access is a SCALAR 128-bit one
// (`.b128`, PTX ISA 8.3+): the PTX memory model makes a scalar access single-copy atomic, whereas a `.v2.b64` vector
// access is formally two 8-byte accesses in unspecified order (ptxas emits the same LDG/STG.E.128.STRONG for both, so
// this costs nothing; it turns "works on today's hardware" into "guaranteed by the model").  A reader therefore never
// sees a new tag next to an old value.  SYS = system scope: the location is written or read by another GPU
// (peer-mapped memory over NVLink; a 16-byte store is one NVLink flit).
template <bool SYS = false>
__device__ __forceinline__ void st_tag(SyncRec* p, double v, u64 g) {
  if constexpr (SYS)
    asm volatile("{\n .reg .b128 q;\n mov.b128 q, {%1, %2};\n st.relaxed.sys.global.b128 [%0], q;\n}" ::"l"(p),
                 "l"(__double_as_longlong(v)), "l"(g)
                 : "memory");
  else
    asm volatile("{\n .reg .b128 q;\n mov.b128 q, {%1, %2};\n st.relaxed.gpu.global.b128 [%0], q;\n}" ::"l"(p),
                 "l"(__double_as_longlong(v)), "l"(g)
                 : "memory");
}

template <bool SYS = false>
__device__ __forceinline__ void ld_tag(const SyncRec* p, double& v, u64& g) {
  long long a;
  if constexpr (SYS)
    asm volatile("{\n .reg .b128 q;\n ld.relaxed.sys.global.b128 q, [%2];\n mov.b128 {%0, %1}, q;\n}" : "=l"(a), "=l"(g) : "l"(p) : "memory");
  else
    asm volatile("{\n .reg .b128 q;\n ld.relaxed.gpu.global.b128 q, [%2];\n mov.b128 {%0, %1}, q;\n}" : "=l"(a), "=l"(g) : "l"(p) : "memory");
  v = __longlong_as_double(a);
}

// ordinary (L1-allocating) load of data that other CTAs wrote during this launch and a fenced barrier made visible
__device__ __forceinline__ double ld_coherent(const double* p) {
  double v;
  asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ u64 global_ns() {
  u64 t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Poll until the element carries `want`.  Gives up after spin_ns (checked every 256 polls) and flags the launch.
template <bool SYS = false>
__device__ __forceinline__ double wait_tag(const SyncRec* p, u64 want, int* fail, u64 spin_ns) {
  double v;
  u64 g;
  ld_tag<SYS>(p, v, g);
  if (g == want) return v;
  const u64 t_begin = global_ns();
  unsigned spins = 0;
  while (true) {
    ld_tag<SYS>(p, v, g);
    if (g == want) return v;
    if ((++spins & 255u) == 0u) {
      if (*(volatile int*)fail) break;  // somebody already gave up: do not wait out the limit again
      if (global_ns() - t_begin > spin_ns) break;
    }
  }
  *fail = 1;
  return v;
}

// ---- grid-wide (and rank-wide) sums ---------------------------------------------------------------------------
// Workers post tagged per-CTA partial sums and later wait for the tagged totals; the reducer CTA (which
// owns no rows) polls all partial sums of a generation (one record per thread), adds them in a fixed order and
// publishes the totals.  Work placed between post and wait overlaps the reduction.  `gen` is the same in
// every CTA (and on every rank) and increases by one per reduction; two record sets alternate on its parity (a
// worker cannot be more than one reduction ahead of the slowest one: it needs a total that includes that
// worker's partial sum; the same argument holds between ranks because the counter continues across launches).
// MULTI: the reducer of every rank stores its rank-local sums into the cross-rank records of ALL ranks (peer
// stores over NVLink), polls its own copy until all ranks have arrived and adds them in rank order: every rank
// gets bit-identical totals, so all ranks take the same branch on the convergence test - an all-reduce without
// a collective call and without leaving the kernel.
__device__ __forceinline__ SyncRec* partial_recs(const PdeArgs& a, u64 gen) { return a.recs + (size_t)(gen & 1ull) * 4 * a.n_workers; }
__device__ __forceinline__ SyncRec* total_recs(const PdeArgs& a, u64 gen) { return a.recs + (size_t)8 * a.n_workers + (gen & 1ull) * 4; }

// FENCE (streaming kernel): the reduction doubles as a grid barrier with memory ordering - everything the CTAs wrote with
// plain stores before post() is visible to plain loads issued after wait() (the cooperative-groups grid.sync pattern:
// bar.sync ; fence ; flag store ... flag load ; fence ; bar.sync - the fence of the polling thread also invalidates the L1).
template <int NV, bool FENCE = false>
__device__ __forceinline__ void post(const double (&v)[NV], const PdeArgs& a, u64 gen, Scratch& sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) sh.warp_part[k * kWarpsPerBlock + warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
    SyncRec* part = partial_recs(a, gen);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = lane < kWarpsPerBlock ? sh.warp_part[k * kWarpsPerBlock + lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) {
        if constexpr (FENCE) __threadfence();
        st_tag(part + (size_t)k * a.n_workers + blockIdx.x, s, gen);
      }
    }
  }
}

// (Measured alternatives, both rejected: (a) no reducer, every CTA polls all 3 x 147 partial sums itself - one L2
// hop fewer on paper, but 147 SMs hammering the same 55 cache lines made the wait 2.2 us instead of 0.3-0.9 us;
// (b) a private copy of the totals per worker CTA to avoid 147 pollers on one line - the 441 extra stores and the
// extra barrier in the reducer cost more than the hot line: PDE stage 51 us instead of 45 us;
// (c) thread-block clusters: a service warp per CTA, CTA sums handed to the cluster leader through distributed shared
// memory + mbarriers, leaders all-to-all through tagged global records, totals written back into the members' shared
// memory - micro-benchmarked at 2.87 us per reduction (8-CTA clusters, 15 leaders) against 2.14 us for this
// reducer-CTA scheme: the two DSMEM hand-offs are not cheaper than the L2 hop they replace, and the chain is longer.)
template <int NV, bool FENCE = false>
__device__ __forceinline__ void wait(double (&v)[NV], const PdeArgs& a, u64 gen, Scratch& sh) {
  if (threadIdx.x < NV) {
    sh.totals[threadIdx.x] = wait_tag(total_recs(a, gen) + threadIdx.x, gen, &sh.fail, a.spin_ns);
    if constexpr (FENCE) __threadfence();
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = sh.totals[k];
  __syncthreads();
}

template <int NV, bool MULTI, bool FENCE = false>
__device__ __forceinline__ void reduce_publish(double (&v)[NV], const PdeArgs& a, u64 gen, Scratch& sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nw = a.n_workers;
  const SyncRec* part = partial_recs(a, gen);
  for (unsigned idx = threadIdx.x; idx < NV * nw; idx += kPdeThreads) {
    sh.red[idx] = wait_tag(part + idx, gen, &sh.fail, a.spin_ns);
    if constexpr (FENCE) __threadfence();  // acquire side of every worker's release; the barrier below makes it cumulative
  }
  __syncthreads();
  if (warp < NV) {
    double s = 0.0;
    for (unsigned i = lane; i < nw; i += 32) s += sh.red[warp * nw + i];
    s = warp_sum(s);
    if constexpr (MULTI) {
      // lane q hands this rank's sum of slot `warp` to rank q, then collects rank q's sum from the local copy
      const size_t slot = ((size_t)(gen & 1ull) * 4 + warp) * a.nranks;
      if (lane < a.nranks) st_tag<true>(a.peer_xrecs[lane] + slot + a.rank, s, gen);
      double mine = 0.0;
      if (lane < a.nranks) mine = wait_tag<true>(a.peer_xrecs[a.rank] + slot + lane, gen, &sh.fail, a.spin_ns);
      s = 0.0;
      for (int q = 0; q < a.nranks; ++q) s += __shfl_sync(0xffffffffu, mine, q);  // rank order: same bits everywhere
    }
    if (lane == 0) {
      if constexpr (FENCE) __threadfence();
      st_tag(total_recs(a, gen) + warp, s, gen);
      sh.totals[warp] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = sh.totals[k];
  __syncthreads();
}

// An owned value that neighbour rows gather: store it (tagged) in this rank's exchange buffer and, for a
// boundary row of a multi-GPU partition, straight into the ghost slot of every neighbour rank that needs it.
template <bool MULTI>
__device__ __forceinline__ void publish(const PdeArgs& a, int which, int64_t row, double v, u64 tag) {
  st_tag(a.tb[which] + row, v, tag);
  if constexpr (MULTI) {
    for (int e = __ldg(a.send_of_row + row); e >= 0;) {
      const SendEnt se = a.send_ents[e];
      st_tag<true>(se.t[which], v, tag);
      e = se.next;
    }
  }
}

// Elements of a gather whose producers are behind: poll ALL of them again, together, until none is left (one L2
// round trip per pass whatever the number of late elements - waiting for them one after the other cost up to 15
// round trips per row when SpMVs follow each other without a reduction in between).
template <bool SYS>
__device__ __forceinline__ unsigned regather(const SyncRec* tv, const int32_t (&c)[kChunk], double (&g)[kChunk], unsigned late, u64 want,
                                             int* fail, u64 spin_ns) {
  u64 t_begin = 0;
  unsigned passes = 0;
  while (late) {
    unsigned still = 0;
#pragma unroll
    for (int u = 0; u < kChunk; ++u) {
      if (late & (1u << u)) {
        u64 t;
        ld_tag<SYS>(tv + c[u], g[u], t);
        still |= (t != want ? 1u : 0u) << u;
      }
    }
    late = still;
    if (late && (++passes & 63u) == 0u) {
      if (t_begin == 0) t_begin = global_ns();
      if (*(volatile int*)fail || global_ns() - t_begin > spin_ns) {
        *fail = 1;
        break;
      }
    }
  }
  return late;
}

// ---- SELL rows ---------------------------------------------------------------------------------------------
// acc[m] = sum_k val(m, k) * g(col(k)) over the `width` entries of a row.  Entries are fetched in unrolled
// batches of kChunk so the index loads, then the gathers, are all in flight together.  TAGGED: the gathered
// vector is an array of {value, generation} pairs and the gather waits until every element carries `want`.
template <int NM, bool TAGGED, bool SYS, bool EARLY_A, bool PARLATE, class ColF, class ValF>
__device__ __forceinline__ void sell_row(int width, ColF col, ValF val, const void* vec, u64 want, int* fail, u64 spin_ns,
                                         double (&acc)[NM]) {
#pragma unroll
  for (int m = 0; m < NM; ++m) acc[m] = 0.0;
  for (int k0 = 0; k0 < width; k0 += kChunk) {
    int32_t c[kChunk];
    double g[kChunk];
#pragma unroll
    for (int u = 0; u < kChunk; ++u) c[u] = (k0 + u < width) ? col(k0 + u) : -1;
    // the matrix entries of the first operand are requested BEFORE the gathers: they do not depend on the column
    // indices, and the tagged loads below are `asm volatile`, which the compiler will not move other loads across -
    // issued after them they would cost a third dependent memory round trip per batch
    // (EARLY_A is off when the entries sit in shared memory: nothing to hide there, and 32 registers saved)
    double av0[kChunk];
    if constexpr (EARLY_A) {
#pragma unroll
      for (int u = 0; u < kChunk; ++u) av0[u] = (k0 + u < width) ? val(0, k0 + u) : 0.0;
    }
    if constexpr (TAGGED) {
      const SyncRec* tv = static_cast<const SyncRec*>(vec);
      unsigned late = 0;
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        g[u] = 0.0;
        if (c[u] >= 0) {
          u64 t;
          ld_tag<SYS>(tv + c[u], g[u], t);
          late |= (t != want ? 1u : 0u) << u;
        }
      }
      if (late) {  // some producers are behind
        if constexpr (PARLATE) {
          regather<SYS>(tv, c, g, late, want, fail, spin_ns);
        } else {  // rare when a reduction separates publish and gather: wait element by element (fewer registers)
#pragma unroll
          for (int u = 0; u < kChunk; ++u)
            if (late & (1u << u)) g[u] = wait_tag<SYS>(tv + c[u], want, fail, spin_ns);
        }
      }
    } else {
      const double* dv = static_cast<const double*>(vec);
#pragma unroll
      for (int u = 0; u < kChunk; ++u) g[u] = c[u] >= 0 ? __ldg(dv + c[u]) : 0.0;
    }
    if constexpr (!EARLY_A) {
#pragma unroll
      for (int u = 0; u < kChunk; ++u) av0[u] = (k0 + u < width) ? val(0, k0 + u) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kChunk; ++u) acc[0] = fma(av0[u], g[u], acc[0]);
#pragma unroll
    for (int m = 1; m < NM; ++m) {
      double av[kChunk];
#pragma unroll
      for (int u = 0; u < kChunk; ++u) av[u] = (k0 + u < width) ? val(m, k0 + u) : 0.0;
#pragma unroll
      for (int u = 0; u < kChunk; ++u) acc[m] = fma(av[u], g[u], acc[m]);
    }
  }
}

// ---- TMA staging of SELL slices (streaming mode) --------------------------------------------------------------
// A slice's entries are two contiguous blocks (values, columns).  In streaming mode shared memory is free, so
// every warp keeps TWO slices in flight with 1-D bulk copies (cp.async.bulk, completion on a per-warp mbarrier)
// while it gathers for a third that it already moved to registers: the HBM stream no longer stalls behind the
// L2 latency of the gathers (measured before: 4.3 TB/s in the SpMV phase with direct loads).
constexpr int kStageValBytes = kChunk * kSlice * 8;
constexpr int kStageBytes = kChunk * kSlice * 12;                       // values + columns of one slice (width <= kChunk)
constexpr int kStagedSmem = kWarpsPerBlock * 2 * kStageBytes + kWarpsPerBlock * 2 * 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// returns a value that depends on the completed wait (ring_at takes it as an ordering token)
__device__ __forceinline__ uint32_t mbar_wait_tok(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
  return ok;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// One warp's view of its two stage buffers; `count` = slices this warp has consumed so far in this launch (the
// stage and the mbarrier parity follow from it, so the pipeline carries over from phase to phase).
struct Stager {
  char* base;        // 2 * kStageBytes of this warp
  uint32_t bar[2];   // its two mbarriers
  unsigned count;
  __device__ __forceinline__ void issue(const PdeArgs& a, const double* vals, int64_t s, unsigned slot) const {  // one lane
    const int64_t beg = __ldg(a.slice_ptr + s);
    const uint32_t w = (uint32_t)((__ldg(a.slice_ptr + s + 1) - beg) / kSlice);
    const unsigned st = slot & 1u;
    mbar_expect_tx(bar[st], w * kSlice * 12u);
    bulk_g2s(smem_u32(base + st * kStageBytes), vals + beg, w * kSlice * 8u, bar[st]);
    bulk_g2s(smem_u32(base + st * kStageBytes + kStageValBytes), a.cols + beg, w * kSlice * 4u, bar[st]);
  }
};

// q_row = sum_k val_k * g(col_k) for the slice `s` of this warp: entries come from the stage buffers (moved to
// registers at once so the buffer can be refilled), the gathered vector is tagged.
template <bool TAGGED, bool SYS>
__device__ __forceinline__ double staged_row(const PdeArgs& a, Stager& S, const double* vals, int64_t s, int64_t s_prefetch, int lane,
                                             const void* vec, u64 want, int* fail) {
  const int64_t beg = __ldg(a.slice_ptr + s);
  const int width = (int)((__ldg(a.slice_ptr + s + 1) - beg) / kSlice);
  const unsigned st = S.count & 1u;
  mbar_wait(S.bar[st], (S.count >> 1) & 1u);
  const double* sA = reinterpret_cast<const double*>(S.base + st * kStageBytes);
  const int32_t* sC = reinterpret_cast<const int32_t*>(S.base + st * kStageBytes + kStageValBytes);
  int32_t c[kChunk];
  double av[kChunk], g[kChunk];
#pragma unroll
  for (int u = 0; u < kChunk; ++u) {
    c[u] = u < width ? sC[u * kSlice + lane] : -1;
    av[u] = u < width ? sA[u * kSlice + lane] : 0.0;
  }
  __syncwarp();
  if (lane == 0 && s_prefetch >= 0) S.issue(a, vals, s_prefetch, S.count + 2);  // s_prefetch: the warp's slice after next, -1: none
  S.count++;
  if constexpr (TAGGED) {
    const SyncRec* tv = static_cast<const SyncRec*>(vec);
    unsigned late = 0;
#pragma unroll
    for (int u = 0; u < kChunk; ++u) {
      g[u] = 0.0;
      if (c[u] >= 0) {
        u64 t;
        ld_tag<SYS>(tv + c[u], g[u], t);
        late |= (t != want ? 1u : 0u) << u;
      }
    }
    if (late) {
#pragma unroll
      for (int u = 0; u < kChunk; ++u)
        if (late & (1u << u)) g[u] = wait_tag<SYS>(tv + c[u], want, fail, a.spin_ns);
    }
  } else {
    const double* dv = static_cast<const double*>(vec);
#pragma unroll
    for (int u = 0; u < kChunk; ++u) g[u] = c[u] >= 0 ? __ldg(dv + c[u]) : 0.0;
  }
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < kChunk; ++u) acc = fma(av[u], g[u], acc);
  return acc;
}

// the same for the streaming kernel's plain search direction: owned columns with ordinary loads, ghost columns (MULTI)
// from the tagged exchange buffer
template <bool MULTI>
__device__ __forceinline__ double staged_row_p(const PdeArgs& a, Stager& S, const double* vals, int64_t s, int64_t s_prefetch, int lane,
                                               const double* p, const SyncRec* ghost_tb, u64 want, int* fail) {
  const int64_t beg = __ldg(a.slice_ptr + s);
  const int width = (int)((__ldg(a.slice_ptr + s + 1) - beg) / kSlice);
  const unsigned st = S.count & 1u;
  mbar_wait(S.bar[st], (S.count >> 1) & 1u);
  const double* sA = reinterpret_cast<const double*>(S.base + st * kStageBytes);
  const int32_t* sC = reinterpret_cast<const int32_t*>(S.base + st * kStageBytes + kStageValBytes);
  int32_t c[kChunk];
  double av[kChunk], g[kChunk];
#pragma unroll
  for (int u = 0; u < kChunk; ++u) {
    c[u] = u < width ? sC[u * kSlice + lane] : -1;
    av[u] = u < width ? sA[u * kSlice + lane] : 0.0;
  }
  __syncwarp();
  if (lane == 0 && s_prefetch >= 0) S.issue(a, vals, s_prefetch, S.count + 2);
  S.count++;
  unsigned ghost = 0;
#pragma unroll
  for (int u = 0; u < kChunk; ++u) {
    g[u] = 0.0;
    if (c[u] >= 0) {
      if (MULTI && c[u] >= a.n_owned)
        ghost |= 1u << u;
      else
        g[u] = ld_coherent(p + c[u]);
    }
  }
  if constexpr (MULTI) {
    if (ghost) {
#pragma unroll
      for (int u = 0; u < kChunk; ++u)
        if (ghost & (1u << u)) g[u] = wait_tag<true>(ghost_tb + c[u], want, fail, a.spin_ns);
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < kChunk; ++u) acc = fma(av[u], g[u], acc);
  return acc;
}

// A thread's view of one of its rows: where the row's entries are (global SELL storage, or the copy of the
// A entries the thread made in shared memory at kernel start when it owns a single row).
struct RowRef {
  int64_t row, beg;
  int width, lane, slot;
};

template <bool MATSMEM, bool PARLATE>
struct MatA {
  const PdeArgs& a;
  const double* sa;    // [kChunk][kPdeThreads] A entries of the thread's row (MATSMEM)
  const int32_t* sc;   // [kChunk][kPdeThreads] their columns
  // gather of the tagged exchange vector; SYS: ghost columns are written by other GPUs
  template <bool SYS>
  __device__ __forceinline__ double apply(const RowRef& r, const void* vec, u64 want, int* fail) const {
    double out[1];
    if constexpr (MATSMEM) {
      sell_row<1, true, SYS, false, PARLATE>(
          r.width, [&](int k) { return sc[k * kPdeThreads + threadIdx.x]; },
          [&](int, int k) { return sa[k * kPdeThreads + threadIdx.x]; }, vec, want, fail, a.spin_ns, out);
    } else {
      sell_row<1, true, SYS, true, PARLATE>(
          r.width, [&](int k) { return __ldg(a.cols + r.beg + (int64_t)k * kSlice + r.lane); },
          [&](int, int k) { return __ldg(a.A + r.beg + (int64_t)k * kSlice + r.lane); }, vec, want, fail, a.spin_ns, out);
    }
    return out[0];
  }
};

// b = B v_ (+ dt * stimulus) and, for a non-zero initial guess x0 = v_, A x0 in the same pass.
template <bool STIM = true>
__device__ __forceinline__ void rhs_row(const PdeArgs& a, const RowRef& r, bool x0_prev, double& bi, double& ax0) {
  int dummy = 0;
  auto col = [&](int k) { return __ldg(a.cols + r.beg + (int64_t)k * kSlice + r.lane); };
  if (x0_prev) {
    double ab[2];
    sell_row<2, false, false, true, false>(
        r.width, col, [&](int m, int k) { return __ldg((m == 0 ? a.B : a.A) + r.beg + (int64_t)k * kSlice + r.lane); }, a.v_prev, 0,
        &dummy, 0, ab);
    bi = ab[0];
    ax0 = ab[1];
  } else {
    double b1[1];
    sell_row<1, false, false, true, false>(
        r.width, col, [&](int, int k) { return __ldg(a.B + r.beg + (int64_t)k * kSlice + r.lane); }, a.v_prev, 0, &dummy, 0, b1);
    bi = b1[0];
    ax0 = 0.0;
  }
  if (STIM && a.has_stim && r.row < a.n_owned) bi = fma(a.dt, __ldg(a.stim_vec + r.row), bi);
}

// ---- stencil dictionary (EXPERIMENTAL) -------------------------------------------------------------------------
// On a structured mesh almost every row repeats one of a few stencils (27 on a box).  Such rows take their entries
// from a table in shared memory - column = row + offset - and never touch the SELL stream; the other rows (pattern
// 255: next to a ghost layer, or a mapped geometry) read their SELL entries as before.  Same entries in the same
// order as the SELL row, so the products and their sum are bit-identical.
struct DictView {
  const double* sA;     // [n_pat][kChunk]
  const double* sB;
  const int32_t* sOff;
  const int32_t* sW;
};

__device__ __forceinline__ DictView dict_load(const PdeArgs& a, double* dyn_smem) {  // all threads of a worker CTA
  double* sA = dyn_smem;
  double* sB = sA + kMaxPat * kChunk;
  int32_t* sOff = reinterpret_cast<int32_t*>(sB + kMaxPat * kChunk);
  int32_t* sW = sOff + kMaxPat * kChunk;
  for (int i = threadIdx.x; i < a.n_pat * kChunk; i += kPdeThreads) {
    sA[i] = __ldg(a.dict_A + i);
    sB[i] = __ldg(a.dict_B + i);
    sOff[i] = __ldg(a.dict_off + i);
  }
  for (int i = threadIdx.x; i < a.n_pat; i += kPdeThreads) sW[i] = __ldg(a.dict_w + i);
  __syncthreads();
  return DictView{sA, sB, sOff, sW};
}

// q_row = sum_k A_k * g(col_k), gathered vector tagged (the CG search direction)
template <bool SYS>
__device__ __forceinline__ double dict_apply(const PdeArgs& a, const DictView& D, const RowRef& r, const void* vec, u64 want, int* fail) {
  const int p = (int)__ldg(a.pat + r.row);
  const bool in_dict = p != 255;
  const int base = in_dict ? p * kChunk : 0;
  const int width = in_dict ? D.sW[p] : r.width;
  double out[1];
  sell_row<1, true, SYS, false, false>(
      width, [&](int k) { return in_dict ? (int32_t)r.row + D.sOff[base + k] : __ldg(a.cols + r.beg + (int64_t)k * kSlice + r.lane); },
      [&](int, int k) { return in_dict ? D.sA[base + k] : __ldg(a.A + r.beg + (int64_t)k * kSlice + r.lane); }, vec, want, fail,
      a.spin_ns, out);
  return out[0];
}

// b = B v_ (+ dt * stimulus) and, for x0 = v_, A x0 in the same pass (the dictionary version of rhs_row)
template <bool STIM = true>
__device__ __forceinline__ void dict_rhs_row(const PdeArgs& a, const DictView& D, const RowRef& r, bool x0_prev, double& bi, double& ax0) {
  int dummy = 0;
  const int p = (int)__ldg(a.pat + r.row);
  const bool in_dict = p != 255;
  const int base = in_dict ? p * kChunk : 0;
  const int width = in_dict ? D.sW[p] : r.width;
  auto col = [&](int k) { return in_dict ? (int32_t)r.row + D.sOff[base + k] : __ldg(a.cols + r.beg + (int64_t)k * kSlice + r.lane); };
  auto val = [&](int m, int k) {
    return in_dict ? (m == 0 ? D.sB : D.sA)[base + k] : __ldg((m == 0 ? a.B : a.A) + r.beg + (int64_t)k * kSlice + r.lane);
  };
  if (x0_prev) {
    double ab[2];
    sell_row<2, false, false, false, false>(width, col, val, a.v_prev, 0, &dummy, 0, ab);
    bi = ab[0];
    ax0 = ab[1];
  } else {
    double b1[1];
    sell_row<1, false, false, false, false>(width, col, val, a.v_prev, 0, &dummy, 0, b1);
    bi = b1[0];
    ax0 = 0.0;
  }
  if (STIM && a.has_stim) bi = fma(a.dt, __ldg(a.stim_vec + r.row), bi);
}

__device__ __forceinline__ double norm_term(int norm_type, double r, double z) {
  // summand of the squared residual norm: preconditioned ||M^-1 r||, unpreconditioned ||r||, natural r.M^-1 r
  return norm_type == MONO_NORM_PRECONDITIONED ? z * z : (norm_type == MONO_NORM_UNPRECONDITIONED ? r * r : r * z);
}

__device__ __forceinline__ int classify(double rnorm, double ttol, double atol, int its, int max_it, bool nan_scalar) {
  if (!(rnorm == rnorm) || nan_scalar) return MONO_KSP_DIVERGED_NAN;
  if (rnorm <= ttol) return rnorm <= atol ? MONO_KSP_CONVERGED_ATOL : MONO_KSP_CONVERGED_RTOL;
  if (its >= max_it) return MONO_KSP_DIVERGED_ITS;
  return 0;
}

__device__ __forceinline__ void write_result(const PdeArgs& a, int its, int reason, double rnorm) {
  if (threadIdx.x == 0) {
    a.res->iterations = its;
    a.res->reason = reason;
    a.res->rnorm = rnorm;
    a.res->total_iterations += its;
    a.res->solves += 1;
  }
}

// Thread-private CG vectors of the rows a thread owns: shared memory (RESIDENT) or global arrays.
enum { VR = 0, VU, VW, VZ, VQ, VS, VP, VN, VX, VD, VY, VE, NVEC };  // streaming mode: VX aliases a.x, VD aliases a.dinv

// Which vectors a kernel keeps per row, and where: KIND 0 = cg (5 vectors), 1 = pipecg/Jacobi (10), 2 = pipecg/Chebyshev
// (12).  Fewer vectors per row = more rows per thread fit the 227 KB of shared memory (cg: 11 rows/thread = 827 k rows
// per GPU stay resident, e.g. each GPU's share of a 3.4 M-dof mesh split over 8).
constexpr int kNvec[3] = {5, 10, 12};
template <int KIND>
__device__ __forceinline__ constexpr int vec_slot(int v) {
  if constexpr (KIND == 0) {
    return v == VR ? 0 : v == VQ ? 1 : v == VP ? 2 : v == VX ? 3 : 4;  // VD
  } else {
    return v;
  }
}

template <bool RESIDENT, int KIND>
struct VecStore {
  double* g[NVEC];
  double* sm;
  int cap;
  __device__ __forceinline__ double ld(int v, const RowRef& r) const {
    if constexpr (RESIDENT) return sm[vec_slot<KIND>(v) * cap + r.slot];
    return __ldcg(g[v] + r.row);
  }
  __device__ __forceinline__ void st(int v, const RowRef& r, double val) const {
    if constexpr (RESIDENT)
      sm[vec_slot<KIND>(v) * cap + r.slot] = val;
    else
      __stcg(g[v] + r.row, val);
  }
};

template <bool RESIDENT, int KIND>
__device__ __forceinline__ VecStore<RESIDENT, KIND> make_store(const PdeArgs& a, double* dyn_smem) {
  VecStore<RESIDENT, KIND> V;
#pragma unroll
  for (int k = 0; k < 8; ++k) V.g[k] = a.work[k];
  V.g[VX] = a.x;
  V.g[VD] = const_cast<double*>(a.dinv);
  V.g[VY] = a.work[8];
  V.g[VE] = a.work[9];
  V.sm = dyn_smem;
  V.cap = a.rows_per_thread * kPdeThreads;
  return V;
}

// rows of this thread: slices warp_global, warp_global + warp_stride, ... (the same in every phase)
template <bool MATSMEM>
__device__ __forceinline__ bool load_row(const PdeArgs& a, RowRef& r, int64_t s, int lane, int slot, int width_cached) {
  r.row = s * kSlice + lane;
  r.lane = lane;
  r.slot = slot;
  if constexpr (MATSMEM) {
    r.beg = 0;
    r.width = width_cached;
  } else {
    r.beg = __ldg(a.slice_ptr + s);
    r.width = (int)((__ldg(a.slice_ptr + s + 1) - r.beg) / kSlice);
  }
  return true;
}

#define OWN_ROWS_BEGIN                                                                                       \
  {                                                                                                          \
    int slot__ = threadIdx.x;                                                                                \
    for (int64_t s__ = warp_global; s__ < a.n_slices; s__ += warp_stride, slot__ += kPdeThreads) {           \
      RowRef r;                                                                                              \
      load_row<MATSMEM>(a, r, s__, lane, slot__, width_cached);
#define OWN_ROWS_END \
  }                  \
  }

// Two rows of the thread per trip (pure vector phases): with the vectors in global memory (streaming mode) the
// loads of both rows are in flight together - one row at a time leaves an SM with ~20 KB outstanding, too little
// to cover HBM latency at full bandwidth.  Rows are still visited in the same order (bit-identical sums).
#define OWN_ROW_PAIRS_BEGIN                                                                                         \
  {                                                                                                                 \
    int slot__ = threadIdx.x;                                                                                       \
    for (int64_t s__ = warp_global; s__ < a.n_slices; s__ += 2 * warp_stride, slot__ += 2 * kPdeThreads) {          \
      RowRef r0, r1;                                                                                                \
      r0.row = s__ * kSlice + lane;                                                                                 \
      r0.lane = lane;                                                                                               \
      r0.slot = slot__;                                                                                             \
      r1 = r0;                                                                                                      \
      r1.row = (s__ + warp_stride) * kSlice + lane;                                                                 \
      r1.slot = slot__ + kPdeThreads;                                                                               \
      const bool on0 = r0.row < a.n_owned;                                                                          \
      const bool on1 = s__ + warp_stride < a.n_slices && r1.row < a.n_owned;
#define OWN_ROW_PAIRS_END \
  }                       \
  }

// Reducer CTA: follows the workers' sequence of reductions and decides convergence exactly as they do (same
// totals, same arithmetic), then reports the result and advances the device-resident generation counter.
__device__ __forceinline__ void finish_reducer(const PdeArgs& a, Scratch& sh, u64 gen, int its, int reason, double rnorm) {
  write_result(a, its, reason, rnorm);
  if (threadIdx.x == 0) {
    a.gen_state[0] = gen;  // every CTA read the old value before its first post, i.e. long before this point
    if (sh.fail) a.res->error = 1;
  }
}

template <bool MULTI>
__device__ void pipecg_reducer(const PdeArgs& a, Scratch& sh, u64 gen) {
  int its = 0, reason = 0;
  double rnorm = 0.0, ttol = 0.0;
  while (true) {
    double acc[4];
    if (its == 0) {  // the first reduction also carries the norm of the right-hand side
      reduce_publish<4, MULTI>(acc, a, gen++, sh);
      ttol = fmax(a.rtol * sqrt(fabs(acc[3])), a.atol);
    } else {
      reduce_publish<3, MULTI>(reinterpret_cast<double(&)[3]>(acc), a, gen++, sh);
    }
    rnorm = sqrt(fabs(acc[2]));
    reason = classify(rnorm, ttol, a.atol, its, a.max_it, !(acc[0] == acc[0]) || !(acc[1] == acc[1]));
    if (sh.fail) reason = MONO_KSP_DIVERGED_NAN;
    if (reason != 0) break;
    ++its;
  }
  finish_reducer(a, sh, gen, its, reason, rnorm);
}

// BARRIER (streaming kernel): the first reduction and one more per iteration (after the new search direction was
// written) are grid barriers with memory ordering, see post().
template <bool MULTI, bool BARRIER = false>
__device__ void cg_reducer(const PdeArgs& a, Scratch& sh, u64 gen) {
  double acc3[3];
  reduce_publish<3, MULTI, BARRIER>(acc3, a, gen++, sh);
  double rnorm = sqrt(fabs(acc3[1]));
  const double ttol = fmax(a.rtol * sqrt(fabs(acc3[2])), a.atol);
  int its = 0;
  int reason = classify(rnorm, ttol, a.atol, 0, a.max_it, false);
  if (sh.fail) reason = MONO_KSP_DIVERGED_NAN;
  while (reason == 0) {
    double pq[1], acc2[2];
    reduce_publish<1, MULTI>(pq, a, gen++, sh);
    reduce_publish<2, MULTI>(acc2, a, gen++, sh);
    ++its;
    rnorm = sqrt(fabs(acc2[1]));
    reason = classify(rnorm, ttol, a.atol, its, a.max_it, !(pq[0] == pq[0]) || pq[0] == 0.0);
    if (sh.fail) reason = MONO_KSP_DIVERGED_NAN;
    if constexpr (BARRIER) {
      if (reason == 0) {
        double one[1];
        reduce_publish<1, MULTI, true>(one, a, gen++, sh);
      }
    }
  }
  finish_reducer(a, sh, gen, its, reason, rnorm);
}

// The A entries of the thread's single row -> shared memory (MATSMEM), once per launch.
template <bool MATSMEM>
__device__ __forceinline__ int stage_matrix(const PdeArgs& a, double* sa, int32_t* sc, int64_t warp_global, int lane) {
  if constexpr (!MATSMEM) return 0;
  int width = 0;
  if (warp_global < a.n_slices) {
    const int64_t beg = __ldg(a.slice_ptr + warp_global);
    width = (int)((__ldg(a.slice_ptr + warp_global + 1) - beg) / kSlice);
    // asynchronous copies (LDGSTS): they complete while the RHS phase runs; stage_wait() precedes the first SpMV.
    // Every thread reads back only what it copied itself, so no barrier is needed.
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
      int32_t* dc = sc + k * kPdeThreads + threadIdx.x;
      double* da = sa + k * kPdeThreads + threadIdx.x;
      if (k < width) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dc)), "l"(a.cols + beg + (int64_t)k * kSlice + lane) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(da)), "l"(a.A + beg + (int64_t)k * kSlice + lane) : "memory");
      } else {
        *dc = -1;
        *da = 0.0;
      }
    }
  }
  return width;
}

template <bool MATSMEM>
__device__ __forceinline__ void stage_wait() {
  if constexpr (MATSMEM) asm volatile("cp.async.wait_all;" ::: "memory");
}

// End of a solve: owned x -> global memory (resident mode keeps it in shared memory until here) and the ghost
// refresh the reference does with state.x.scatter_forward() (base_model.py:242).  MULTI: boundary values go
// straight into the neighbours' landing zones tagged with this launch's final generation, and this rank's
// ghosts are copied out of its own landing zone as they arrive - the next kernel of this rank (the cell-model
// update, which runs on ghosts too) starts only after every neighbour has finished reading this rank's
// exchange buffers, which is what makes their reuse by the next launch safe.
#define FINISH_SOLVE(XTAG)                                                                                   \
  {                                                                                                          \
    const u64 xtag__ = (XTAG);                                                                               \
    OWN_ROWS_BEGIN                                                                                           \
      if (r.row < a.n_owned) {                                                                               \
        const double xi = V.ld(VX, r);                                                                       \
        if constexpr (RESIDENT) a.x[r.row] = xi;                                                             \
        if constexpr (MULTI) {                                                                               \
          for (int e = __ldg(a.send_of_row + r.row); e >= 0;) {                                              \
            const SendEnt se = a.send_ents[e];                                                               \
            st_tag<true>(se.xg, xi, xtag__);                                                                 \
            e = se.next;                                                                                     \
          }                                                                                                  \
        }                                                                                                    \
      }                                                                                                      \
    OWN_ROWS_END                                                                                             \
    if constexpr (MULTI) {                                                                                   \
      const int64_t n_ghost__ = a.n_local - a.n_owned;                                                       \
      for (int64_t g = (int64_t)blockIdx.x * kPdeThreads + threadIdx.x; g < n_ghost__;                       \
           g += (int64_t)a.n_workers * kPdeThreads)                                                          \
        a.x[a.n_owned + g] = wait_tag<true>(a.xg + g, xtag__, &sh.fail, a.spin_ns);                          \
    }                                                                                                        \
    if (sh.fail && threadIdx.x == 0) a.res->error = 1;                                                       \
    if (blockIdx.x == 0 && threadIdx.x == 0) a.gen_state[1] = vtag; /* every CTA read it at its start */      \
  }

// ---- KSPCG (PETSc semantics): two reductions per iteration ------------------------------------------------
//   q = A p ; alpha = (r,z)/(p,q) ; x += alpha p ; r -= alpha q ; z = M^-1 r ; beta = (r,z)_new/(r,z) ; p = z + beta p
template <bool RESIDENT, bool MATSMEM, bool MULTI, bool DICT = false>
__global__ void __launch_bounds__(kPdeThreads, 1) pde_cg_kernel(const PdeArgs a) {
  static_assert(!DICT || (!RESIDENT && !MATSMEM), "the stencil dictionary serves the streaming mode");
  extern __shared__ double dyn_smem[];
  __shared__ Scratch sh;
  if (threadIdx.x == 0) sh.fail = 0;
  __syncthreads();
  u64 gen = a.gen_state[0];   // reduction generations (the same sequence in every CTA and on every rank)
  u64 vtag = a.gen_state[1] + 1;  // tag of the exchanged vector published last: unique per write, continues across launches
  if (blockIdx.x == a.n_workers) {
    cg_reducer<MULTI>(a, sh, gen);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)(threadIdx.x >> 5) * a.n_workers + blockIdx.x;
  const int64_t warp_stride = (int64_t)a.n_workers * kWarpsPerBlock;
  const bool x0_prev = a.x0_mode == MONO_X0_PREVIOUS;
  const VecStore<RESIDENT, 0> V = make_store<RESIDENT, 0>(a, dyn_smem);
  double* sa = dyn_smem + (size_t)kNvec[0] * V.cap;
  int32_t* sc = reinterpret_cast<int32_t*>(sa + kChunk * kPdeThreads);
  const int width_cached = stage_matrix<MATSMEM>(a, sa, sc, warp_global, lane);
  const MatA<MATSMEM, false> Aop{a, sa, sc};
  Stager S{};
  [[maybe_unused]] DictView D{};
  if constexpr (DICT) D = dict_load(a, dyn_smem);  // (the dictionary takes the place of the stage buffers: a.staged is 0)
  if constexpr (!RESIDENT && !DICT) {
    if (a.staged) {  // (block-uniform) per-warp stage buffers + mbarriers in the otherwise unused dynamic shared memory
      const int warp = threadIdx.x >> 5;
      char* smem = reinterpret_cast<char*>(dyn_smem);
      S.base = smem + warp * 2 * kStageBytes;
      S.bar[0] = smem_u32(smem + kWarpsPerBlock * 2 * kStageBytes + warp * 16);
      S.bar[1] = S.bar[0] + 8;
      S.count = 0;
      if (lane == 0) {
        mbar_init(S.bar[0], 1);
        mbar_init(S.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncthreads();
    }
  }
  int nstamp = 0;
  stamp(a, nstamp);

  // ---- K2 + initial residual: r = b - A x0, z = D^-1 r, p = z ---------------------------------------------
  double acc3[3] = {0.0, 0.0, 0.0};  // r.z, norm^2 of r, norm^2 of b (chosen norm)
  int cur = 0;
  auto rhs_finish = [&](const RowRef& r, double bi, double ax0) {
    const double di = __ldg(a.dinv + r.row);
    if constexpr (RESIDENT) V.st(VD, r, di);
    const double ri = x0_prev ? bi - ax0 : bi;
    const double zi = di * ri;
    V.st(VX, r, x0_prev ? __ldg(a.v_prev + r.row) : 0.0);
    V.st(VR, r, ri);
    V.st(VP, r, zi);
    publish<MULTI>(a, cur, r.row, zi, vtag);
    acc3[0] = fma(ri, zi, acc3[0]);
    acc3[1] += norm_term(a.norm_type, ri, zi);
    acc3[2] += norm_term(a.norm_type, bi, di * bi);
  };
  bool rhs_staged = false;
  if constexpr (DICT) {
    rhs_staged = true;
    OWN_ROWS_BEGIN
      if (r.row < a.n_owned) {
        double bi, ax0;
        dict_rhs_row(a, D, r, x0_prev, bi, ax0);
        rhs_finish(r, bi, ax0);
      }
    OWN_ROWS_END
  }
  if constexpr (!RESIDENT && !DICT) {
    if (a.staged && !x0_prev) {  // b = B v_ with the B slices arriving through the stage buffers
      rhs_staged = true;
      if (lane == 0) {
        if (warp_global < a.n_slices) S.issue(a, a.B, warp_global, S.count);
        if (warp_global + warp_stride < a.n_slices) S.issue(a, a.B, warp_global + warp_stride, S.count + 1);
      }
      for (int64_t s = warp_global; s < a.n_slices; s += warp_stride) {
        RowRef r;
        r.row = s * kSlice + lane;
        r.lane = lane;
        r.slot = 0;
        const bool on = r.row < a.n_owned;
        const double sv = (on && a.has_stim) ? __ldg(a.stim_vec + r.row) : 0.0;
        double bi = staged_row<false, false>(a, S, a.B, s, s + 2 * warp_stride < a.n_slices ? s + 2 * warp_stride : -1, lane, a.v_prev, 0, &sh.fail);
        if (on) {
          if (a.has_stim) bi = fma(a.dt, sv, bi);
          rhs_finish(r, bi, 0.0);
        }
      }
    }
  }
  if (!rhs_staged) {
    OWN_ROWS_BEGIN
      RowRef g = r;
      if constexpr (MATSMEM) {
        g.beg = __ldg(a.slice_ptr + s__);
      }
      double bi, ax0;
      rhs_row(a, g, x0_prev, bi, ax0);
      if (r.row < a.n_owned) rhs_finish(r, bi, ax0);
    OWN_ROWS_END
  }
  stamp(a, nstamp);
  post<3>(acc3, a, gen, sh);
  wait<3>(acc3, a, gen++, sh);
  stamp(a, nstamp);
  double rz = acc3[0];
  double rnorm = sqrt(fabs(acc3[1]));
  const double ttol = fmax(a.rtol * sqrt(fabs(acc3[2])), a.atol);
  int its = 0;
  int reason = classify(rnorm, ttol, a.atol, 0, a.max_it, false);
  if (sh.fail) reason = MONO_KSP_DIVERGED_NAN;
  stage_wait<MATSMEM>();
  while (reason == 0) {
    // ---- K4a: q = A p (gathers wait on the tag of each element), p.q ----------------------------------------
    double pq[1] = {0.0};
    bool staged_done = false;
    if constexpr (DICT) {
      staged_done = true;
      OWN_ROWS_BEGIN
        if (r.row < a.n_owned) {
          const double pi = __ldcg(a.work[VP] + r.row);  // requested before the gathers
          const double qi = dict_apply<MULTI>(a, D, r, a.tb[cur], vtag, &sh.fail);
          __stcg(a.work[VQ] + r.row, qi);
          pq[0] = fma(pi, qi, pq[0]);
        }
      OWN_ROWS_END
    }
    if constexpr (!RESIDENT && !DICT) {
      if (a.staged) {
        staged_done = true;
        if (lane == 0) {
          if (warp_global < a.n_slices) S.issue(a, a.A, warp_global, S.count);
          if (warp_global + warp_stride < a.n_slices) S.issue(a, a.A, warp_global + warp_stride, S.count + 1);
        }
        for (int64_t s = warp_global; s < a.n_slices; s += warp_stride) {
          const int64_t row = s * kSlice + lane;
          const double pi = row < a.n_owned ? __ldcg(a.work[VP] + row) : 0.0;  // requested before the gathers
          const double qi = staged_row<true, MULTI>(a, S, a.A, s, s + 2 * warp_stride < a.n_slices ? s + 2 * warp_stride : -1, lane, a.tb[cur], vtag, &sh.fail);
          if (row < a.n_owned) {
            __stcg(a.work[VQ] + row, qi);
            pq[0] = fma(pi, qi, pq[0]);
          }
        }
      }
    }
    if (!staged_done) {
      OWN_ROWS_BEGIN
        const double qi = Aop.template apply<MULTI>(r, a.tb[cur], vtag, &sh.fail);
        if (r.row < a.n_owned) {
          V.st(VQ, r, qi);
          pq[0] = fma(V.ld(VP, r), qi, pq[0]);
        }
      OWN_ROWS_END
    }
    stamp(a, nstamp);
    post<1>(pq, a, gen, sh);
    wait<1>(pq, a, gen++, sh);
    stamp(a, nstamp);
    const double alpha = rz / pq[0];
    // ---- K4b: x += alpha p ; r -= alpha q ; z = D^-1 r ; r.z and the residual norm (own rows) ------------------
    double acc2[2] = {0.0, 0.0};
    {
      struct In { double q, r, d, p, x; };
      auto load = [&](const RowRef& r) { return In{V.ld(VQ, r), V.ld(VR, r), V.ld(VD, r), V.ld(VP, r), V.ld(VX, r)}; };
      auto finish = [&](const RowRef& r, const In& v) {
        const double ri = fma(-alpha, v.q, v.r);
        const double zi = v.d * ri;
        V.st(VX, r, fma(alpha, v.p, v.x));
        V.st(VR, r, ri);
        acc2[0] = fma(ri, zi, acc2[0]);
        acc2[1] += norm_term(a.norm_type, ri, zi);
      };
      OWN_ROW_PAIRS_BEGIN
        In v0{}, v1{};
        if (on0) v0 = load(r0);
        if (on1) v1 = load(r1);
        if (on0) finish(r0, v0);
        if (on1) finish(r1, v1);
      OWN_ROW_PAIRS_END
    }
    stamp(a, nstamp);
    post<2>(acc2, a, gen, sh);
    wait<2>(acc2, a, gen++, sh);
    stamp(a, nstamp);
    ++its;
    rnorm = sqrt(fabs(acc2[1]));
    const double beta = acc2[0] / rz;
    rz = acc2[0];
    reason = classify(rnorm, ttol, a.atol, its, a.max_it, !(pq[0] == pq[0]) || pq[0] == 0.0);
    if (sh.fail) reason = MONO_KSP_DIVERGED_NAN;
    if (reason != 0) break;
    // ---- p = z + beta p (own rows), published with the next generation tag ------------------------------------
    ++vtag;
    cur ^= 1;
    {
      struct In { double p, d, r; };
      auto load = [&](const RowRef& r) { return In{V.ld(VP, r), V.ld(VD, r), V.ld(VR, r)}; };
      auto finish = [&](const RowRef& r, const In& v) {
        const double pi = fma(beta, v.p, v.d * v.r);
        V.st(VP, r, pi);
        publish<MULTI>(a, cur, r.row, pi, vtag);
      };
      OWN_ROW_PAIRS_BEGIN
        In v0{}, v1{};
        if (on0) v0 = load(r0);
        if (on1) v1 = load(r1);
        if (on0) finish(r0, v0);
        if (on1) finish(r1, v1);
      OWN_ROW_PAIRS_END
    }
    stamp(a, nstamp);
  }
  FINISH_SOLVE(gen)
}

// ---- KSPCG for meshes that stream from HBM: plain vectors, one memory-ordered grid barrier per iteration ------------
// Same iterates as pde_cg_kernel.  What is different is how data crosses CTAs, because the trade-off flips with size:
// a resident mesh is latency bound (a fenced barrier costs microseconds, so values travel with tags and nothing ever
// waits for a barrier), a streamed mesh is bandwidth bound (a phase takes 50-500 us, a fenced barrier ~3 us), and there
// the tagged exchange costs 16 B per element on every gather plus a second copy of the search direction.  So here
//   * the search direction p is ONE plain fp64 vector; the reduction that follows its update is a grid barrier with
//     release/acquire ordering (post/wait/reduce_publish with FENCE), after which the SpMV gathers it with ordinary
//     L1-cached loads (the acquire fence of the polling thread invalidated the L1) - 8 B per gathered element, and the
//     elements of neighbouring rows come out of the L1 instead of the L2;
//   * only GHOST columns (multi-GPU) still use the tagged exchange buffers: peers store them straight into this rank's
//     ghost slots and the gather polls the tag of exactly the ghost elements it needs, so no cross-GPU fence exists;
//   * every CTA owns one contiguous run of SELL slices (warp w takes slices begin + w, + 16, ...): the y-neighbours of
//     a row were gathered by the same SM a few trips earlier, so they hit its L1; 147 sequential streams keep the HBM
//     pages open;
//   * vector phases take four rows per thread and trip, dictionary SpMV rows two, so that a 512-thread CTA keeps
//     ~80 KB of loads in flight (Little: 43 GB/s per SM x ~1 us).
// Per iteration and row (dictionary rows): SpMV 8 (p, once per line through L1/L2) + 8 (q) + 1 (pattern), update
// 40 + 16, direction 24 + 8 = 105 B against 137 B + 240 B of L2 gather traffic for the tagged kernel.
// gather of the search direction for one SELL/dictionary entry list: owned columns from the plain vector, ghost columns
// (MULTI) from the tagged exchange buffer the owner rank stores into
template <bool MULTI, class ColF, class ValF>
__device__ __forceinline__ double row_times_p(const PdeArgs& a, int width, ColF col, ValF val, const double* p, const SyncRec* ghost_tb,
                                              u64 want, int* fail) {
  double acc = 0.0;
  for (int k0 = 0; k0 < width; k0 += kChunk) {
    int32_t c[kChunk];
    double g[kChunk];
    unsigned ghost = 0;
#pragma unroll
    for (int u = 0; u < kChunk; ++u) c[u] = (k0 + u < width) ? col(k0 + u) : -1;
#pragma unroll
    for (int u = 0; u < kChunk; ++u) {
      g[u] = 0.0;
      if (c[u] >= 0) {
        if (MULTI && c[u] >= a.n_owned)
          ghost |= 1u << u;
        else
          g[u] = ld_coherent(p + c[u]);
      }
    }
    if constexpr (MULTI) {
      if (ghost) {
#pragma unroll
        for (int u = 0; u < kChunk; ++u)
          if (ghost & (1u << u)) g[u] = wait_tag<true>(ghost_tb + c[u], want, fail, a.spin_ns);
      }
    }
#pragma unroll
    for (int u = 0; u < kChunk; ++u) acc = fma((k0 + u < width) ? val(k0 + u) : 0.0, g[u], acc);
  }
  return acc;
}

// ---- shared-memory rings for the dictionary rows (MODE 2) ----------------------------------------------------------
// On a structured mesh the column offsets of all stencils fall into a few clusters (the planes of the mesh: for the
// Kuhn-split box z-1, z, z+1).  A CTA sweeps its contiguous run of rows in tiles of 512 (one row per thread); for every
// cluster a ring buffer in shared memory follows the sweep: before a tile is computed, one thread has asked the TMA
// (cp.async.bulk, completion on an mbarrier) for the part of [tile_begin + lo_c, tile_end + hi_c] of the gathered vector
// that is not in the ring yet - every element of the vector travels L2 -> shared memory ONCE per cluster, in bulk, two
// tiles ahead, with no register or L1 involvement - and the 15 gathers of a row become conflict-free shared-memory loads
// ring_c[(row + off) mod cap].  Rows outside the dictionary (next to a ghost layer; multi-GPU) are skipped here and come
// from a compact row-major side table, thread per row.
constexpr int kModeSell = 0, kModeDictL1 = 1, kModeRing = 2;

struct RingView {
  double* buf;         // [ncl][cap]
  uint32_t buf_u32;    // its shared-memory address
  uint32_t full0;      // kRingStages mbarriers "the ring data of this tile has landed" (transaction count) ...
  uint32_t empty0;     // ... and kRingStages mbarriers "all 16 warps are done with this tile" (one arrival per warp)
  uint32_t stage_u32;  // kRingStages x kRingStageBytes: pattern bytes + stimulus flags of the tiles in flight
  const uint8_t* stage_buf;
  int cap_log2;
  int nst;             // stages in use = tiles in flight + 1
  int stage;           // stage of the next tile to consume ...
  unsigned parity;     // ... and the parity its mbarriers complete with (both continue from sweep to sweep)
  unsigned tiles;      // tiles consumed so far in this launch
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Requests for the ring data of tile j of a sweep, in 256-element chunks (256-aligned, ring capacity a multiple of 256: a
// chunk never wraps).  Stateless: what tile j adds to cluster c is [ceil256(need(j-1)), ceil256(need(j))) with
// need(j) = tile_end(j) + hi_c (tile 0 starts at floor256(row_b + lo_c)), so ANY thread can issue any chunk.  That matters:
// a cp.async.bulk costs its issuing thread several hundred cycles (measured: with one requesting thread the whole sweep
// ran at exactly the issue rate of that thread, ~3000 cycles per tile whatever the prefetch depth), so the chunks of a tile
// are dealt over the 16 warps - lane 0 of warp w issues chunk k of cluster c when (5 c + k) mod 16 == w and then arrives on
// the tile's full barrier (16 arrivals) with the bytes it asked for.
__device__ __forceinline__ void ring_issue_share(const PdeArgs& a, const RingView& R, const double* src, int64_t n_even, int stage,
                                                 int64_t row_b, int64_t row_e, int64_t j, int warp, bool wait_empty, unsigned empty_parity) {
  const uint32_t bar = R.full0 + (uint32_t)stage * 8u;
  const int64_t mask = (1ll << R.cap_log2) - 1;
  const int64_t tile_hi = min(row_b + (j + 1) * kRingTile, row_e), prev_hi = min(row_b + j * kRingTile, row_e);
  uint32_t bytes = 0;
  bool waited = !wait_empty;
#pragma unroll
  for (int c = 0; c < kMaxClusters; ++c) {
    if (c < a.ring_ncl) {
      const int64_t end = (min(tile_hi + a.ring_hi[c], n_even) + kRingChunk - 1) & ~(int64_t)(kRingChunk - 1);
      const int64_t beg = j == 0 ? ((row_b + a.ring_lo[c]) >> kRingChunkLog2) << kRingChunkLog2
                                 : (min(prev_hi + a.ring_hi[c], n_even) + kRingChunk - 1) & ~(int64_t)(kRingChunk - 1);
      const int64_t nch = end > beg ? (end - beg) >> kRingChunkLog2 : 0;
      for (int64_t k = (a.dbg & 8) ? (warp == 0 ? 0 : nch) : ((warp - 5 * c) & 15); k < nch; k += (a.dbg & 8) ? 1 : 16) {
        const int64_t lo = max(beg + k * kRingChunk, (int64_t)0), hi = min(beg + (k + 1) * kRingChunk, n_even);
        if (hi > lo && !((a.dbg & 1) && src != a.v_prev)) {
          if (!waited) {  // the ring space this chunk overwrites was read by tile (j - stages): wait until all warps are done with it
            mbar_wait(R.empty0 + (uint32_t)stage * 8u, empty_parity);
            waited = true;
          }
          bulk_g2s(R.buf_u32 + (uint32_t)((((int64_t)c << R.cap_log2) + (lo & mask)) * 8), src + lo, (uint32_t)(hi - lo) * 8u, bar);
          bytes += (uint32_t)(hi - lo) * 8u;
        }
      }
    }
  }
  if (warp == kWarpsPerBlock - 1) {  // the tile's pattern bytes and stimulus flags (buffers of this stage: same hand-shake)
    if (!waited) mbar_wait(R.empty0 + (uint32_t)stage * 8u, empty_parity);
    const int64_t tile_lo = row_b + j * kRingTile;  // (a multiple of the tile: the CTA's run of slices starts on a multiple of 32)
    bulk_g2s(R.stage_u32 + (uint32_t)stage * kRingStageBytes, a.pat + tile_lo, kRingTile, bar);
    bytes += kRingTile;
    if (a.has_stim && src == a.v_prev) {  // (only the right-hand side needs the stimulus flags, and only while a stimulus is on)
      bulk_g2s(R.stage_u32 + (uint32_t)stage * kRingStageBytes + kRingTile, a.slice_stim + tile_lo / kSlice, kRingTile / kSlice, bar);
      bytes += kRingTile / kSlice;
    }
  }
  mbar_expect_tx(bar, bytes);  // (arrives too: the phase cannot complete before all 16 warps have, whatever has landed already)
}

// Sweep of the CTA's rows [row_b, row_e) in tiles, the rings following `src`; f(row, tok, pid, stim_flag) runs once per row
// after the data of its tile has landed: ring elements, the row's pattern byte and its slice's stimulus flag all arrive
// through bulk copies on the tile's full barrier, so the consumer path has NO global load in front of its compute (a
// per-tile dependent global load - even one requested several tiles ahead into a register queue, whose shifting moves
// wait for the loads - costs a memory latency per tile: measured, 29 % of all stall samples on one branch).
// The 16 warps are NOT synchronised per tile: a warp takes the next tile as soon as ITS data is there (full barrier) and
// says so when it is done (empty barrier); a warp that has a chunk to request waits for the tile whose ring space the
// chunk overwrites, `depth + 1` tiles back.
template <class RowF>
__device__ __forceinline__ void ring_sweep(const PdeArgs& a, RingView& R, const double* src, int64_t n_src, int64_t row_b, int64_t row_e,
                                           RowF f) {
  const int64_t ntiles = (row_e - row_b + kRingTile - 1) / kRingTile;
  const int64_t n_even = (n_src + 1) & ~1ll;  // (the source arrays are padded: reading one element past the end is fine)
  const int depth = R.nst - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int pstage = R.stage;            // requests (lane 0 of every warp): stage, parity and launch-wide index of the next tile to request
  unsigned pparity = R.parity, ptile = R.tiles;
  // Steady state (whole tiles after the first): every cluster advances by exactly one chunk per tile, chunk c =
  // [E0_c - 1024 + 1024 j, + 1024) with E0_c the chunk-aligned end of what tile 0 needs; warp c requests it and the last
  // warp the tile's pattern bytes - a compare or two and one bulk copy per warp and tile.  The request path must be cheap
  // and the copies FEW: measured, a sweep costs ~0.2-0.4 us per bulk copy and SM whatever its size (8 copies of <= 2 KB per
  // 512 rows: 1.6 us per tile even with compute and stores switched off), hence 1024-row tiles and 8 KB chunks.
  // Tile 0 and a partial last tile take the general path.
  const int cl = warp;
  const bool has_chunk = warp < a.ring_ncl;
  const int64_t chunk0 = has_chunk ? ((min(row_b + kRingTile, row_e) + a.ring_hi[cl] + kRingChunk - 1) & ~(int64_t)(kRingChunk - 1)) -
                                         kRingTile
                                   : 0;
  const uint32_t ring_c = R.buf_u32 + (uint32_t)(((int64_t)cl << R.cap_log2) * 8);
  const uint32_t pos_mask = (1u << R.cap_log2) - 1u;
  const int64_t full_tiles = (row_e - row_b) / kRingTile;
  auto request = [&](int64_t j) {
    if (j >= 1 && j < full_tiles) {
      const uint32_t bar = R.full0 + (uint32_t)pstage * 8u;
      uint32_t bytes = 0;
      const bool wait_empty = ptile >= (unsigned)R.nst;
      if (has_chunk) {
        const int64_t lo = chunk0 + j * kRingTile;
        if (lo >= 0 && lo < n_even) {
          if (wait_empty) mbar_wait(R.empty0 + (uint32_t)pstage * 8u, pparity ^ 1u);
          bytes = (uint32_t)min((int64_t)kRingChunk, n_even - lo) * 8u;
          bulk_g2s(ring_c + (((uint32_t)lo & pos_mask) << 3), src + lo, bytes, bar);
        }
      } else if (warp == kWarpsPerBlock - 1) {
        if (wait_empty) mbar_wait(R.empty0 + (uint32_t)pstage * 8u, pparity ^ 1u);
        const int64_t tile_lo = row_b + j * kRingTile;
        bulk_g2s(R.stage_u32 + (uint32_t)pstage * kRingStageBytes, a.pat + tile_lo, kRingTile, bar);
        bytes = kRingTile;
        if (a.has_stim && src == a.v_prev) {
          bulk_g2s(R.stage_u32 + (uint32_t)pstage * kRingStageBytes + kRingTile, a.slice_stim + tile_lo / kSlice, kRingTile / kSlice, bar);
          bytes += kRingTile / kSlice;
        }
      }
      mbar_expect_tx(bar, bytes);
    } else {
      ring_issue_share(a, R, src, n_even, pstage, row_b, row_e, j, warp, ptile >= (unsigned)R.nst, pparity ^ 1u);
    }
    ++ptile;
    if (++pstage == R.nst) {
      pstage = 0;
      pparity ^= 1u;
    }
  };
  if (lane == 0) {
    asm volatile("fence.proxy.async.global;" ::: "memory");  // the vector was written with ordinary stores (other CTAs, before the barrier)
    for (int64_t t = 0; t < min((int64_t)depth, ntiles); ++t) request(t);
  }
  for (int64_t t = 0; t < ntiles; ++t) {
    const int64_t row = row_b + t * kRingTile + threadIdx.x;
    if (lane == 0 && t + depth < ntiles) request(t + depth);
    __syncwarp();
    const unsigned tok = mbar_wait_tok(R.full0 + (uint32_t)R.stage * 8u, R.parity);
    const uint8_t* st = R.stage_buf + R.stage * kRingStageBytes;
#pragma unroll
    for (int k = 0; k < kRingRows; ++k)  // thread tid owns rows tile + tid, tile + 512 + tid (as in the vector phases)
      f(row + k * kPdeThreads, tok, (int)st[k * kPdeThreads + threadIdx.x], (int)st[kRingTile + k * kWarpsPerBlock + warp]);
    __syncwarp();
    if (lane == 0) mbar_arrive(R.empty0 + (uint32_t)R.stage * 8u);
    ++R.tiles;
    if (++R.stage == R.nst) {
      R.stage = 0;
      R.parity ^= 1u;
    }
  }
  __syncthreads();  // the rings may be re-filled (next sweep) only after the last tile was read
}

// the values of one stencil in registers (reloaded from the shared-memory dictionary when the row's pattern changes); the
// ring positions come from a table of {offset, byte offset of the cluster's ring} pairs, two per (broadcast) 16-byte load:
// address = cluster + ((row + offset) & mask) * 8 - three integer instructions per entry.  All table words, then all ring
// loads, are issued before the multiply-add chain starts (the chain keeps the entry order of the SELL row: same bits).
struct PatRegs {
  int id;
  double v[kChunk];
};

__device__ __forceinline__ void pat_load(PatRegs& P, int pid, const double* sV) {
  P.id = pid;
#pragma unroll
  for (int u = 0; u < kChunk; ++u) P.v[u] = sV[pid * kChunk + u];
}

// (not `asm volatile`: volatile asm statements keep their order, which would serialise the 16 loads of a row behind each
// other's address arithmetic; `tok` - a value the wait on the tile's full barrier produced - is what keeps the load after it)
__device__ __forceinline__ double ring_at(uint32_t ring_u32, uint2 e, unsigned r, unsigned mask, unsigned tok) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(ring_u32 + e.y + (((r + e.x) & mask) << 3)), "r"(tok));
  return v;
}

__device__ __forceinline__ double ring_row(const PatRegs& P, const uint4* sEnc, uint32_t ring_u32, int64_t row, unsigned mask, unsigned tok) {
  const unsigned r = (unsigned)row;
  uint4 e[kChunk / 2];
#pragma unroll
  for (int u2 = 0; u2 < kChunk / 2; ++u2) e[u2] = sEnc[P.id * (kChunk / 2) + u2];
  double g[kChunk];
#pragma unroll
  for (int u2 = 0; u2 < kChunk / 2; ++u2) {
    g[2 * u2] = ring_at(ring_u32, make_uint2(e[u2].x, e[u2].y), r, mask, tok);
    g[2 * u2 + 1] = ring_at(ring_u32, make_uint2(e[u2].z, e[u2].w), r, mask, tok);
  }
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < kChunk; ++u) acc = fma(P.v[u], g[u], acc);
  return acc;
}

template <bool MULTI, int MODE>
__global__ void __launch_bounds__(kPdeThreads, 1) pde_cg_stream_kernel(const PdeArgs a) {
  constexpr bool DICT = MODE != kModeSell;
  extern __shared__ double dyn_smem[];
  __shared__ Scratch sh;
  if (threadIdx.x == 0) sh.fail = 0;
  __syncthreads();
  u64 gen = a.gen_state[0];       // reduction generations (the same sequence in every CTA and on every rank)
  u64 vtag = a.gen_state[1] + 1;  // tag of the ghost values published last: unique per write, continues across launches
  if (blockIdx.x == a.n_workers) {
    cg_reducer<MULTI, true>(a, sh, gen);
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // this CTA's contiguous run of slices; warp w owns slices s_first, s_first + 16, ... in EVERY phase, i.e. thread tid owns
  // rows row_b + tid, row_b + 512 + tid, ...
  // (a whole number of ring tiles per CTA, so that every tile starts on a multiple of the tile size; the same split in
  // every mode, so that the modes add their dot products in the same order and agree bit for bit)
  int64_t per_cta = (a.n_slices + a.n_workers - 1) / a.n_workers;
  per_cta = (per_cta + kRingTile / kSlice - 1) / (kRingTile / kSlice) * (kRingTile / kSlice);
  const int64_t s_begin = min((int64_t)blockIdx.x * per_cta, a.n_slices);
  const int64_t s_end = min(s_begin + per_cta, a.n_slices);
  const int64_t s_first = s_begin + warp;
  const int64_t row_b = s_begin * kSlice, row_e = min(s_end * kSlice, a.n_owned);
  constexpr int64_t W = kWarpsPerBlock;
  const bool x0_prev = a.x0_mode == MONO_X0_PREVIOUS;
  double* const vr = a.work[VR];
  double* const vq = a.work[VQ];
  double* const vp = a.work[VP];   // the search direction: plain, gathered by every CTA after the barrier
  double* const vx = a.x;
  [[maybe_unused]] DictView D{};
  [[maybe_unused]] const uint4* sEnc = nullptr;
  [[maybe_unused]] const double* sDinv = nullptr;
  [[maybe_unused]] RingView R{};
  Stager S{};
  if constexpr (MODE == kModeDictL1) {
    D = dict_load(a, dyn_smem);
  } else if constexpr (MODE == kModeRing) {
    double* sA = dyn_smem;
    double* sB = sA + kMaxPat * kChunk;
    uint2* sEncw = reinterpret_cast<uint2*>(sB + kMaxPat * kChunk);
    double* sDinvw = reinterpret_cast<double*>(sEncw + kMaxPat * kChunk);
    int32_t* sW = reinterpret_cast<int32_t*>(sDinvw + kMaxPat);
    for (int i = threadIdx.x; i < kMaxPat * kChunk; i += kPdeThreads) {
      const bool in = i < a.n_pat * kChunk;
      sA[i] = in ? __ldg(a.dict_A + i) : 0.0;
      sB[i] = in ? __ldg(a.dict_B + i) : 0.0;
      sEncw[i] = in ? make_uint2((unsigned)__ldg(a.dict_off + i), ((unsigned)__ldg(a.dict_cl + i) << a.ring_cap_log2) * 8u)
                    : make_uint2(0u, (unsigned)(a.ring_cl0 << a.ring_cap_log2) * 8u);
    }
    for (int i = threadIdx.x; i < kMaxPat; i += kPdeThreads) {
      sW[i] = i < a.n_pat ? __ldg(a.dict_w + i) : 0;
      sDinvw[i] = i < a.n_pat ? __ldg(a.dict_dinv + i) : 0.0;
    }
    D = DictView{sA, sB, nullptr, sW};
    sEnc = reinterpret_cast<const uint4*>(sEncw);
    sDinv = sDinvw;
    char* ring_base = reinterpret_cast<char*>(dyn_smem) + ((kRingTableSmem + 127) / 128) * 128;
    R.buf = reinterpret_cast<double*>(ring_base);
    R.buf_u32 = smem_u32(ring_base);
    R.cap_log2 = a.ring_cap_log2;
    char* after = ring_base + ((size_t)a.ring_ncl << a.ring_cap_log2) * 8;
    R.full0 = smem_u32(after);
    R.empty0 = R.full0 + 8u * kRingStages;
    R.stage_buf = reinterpret_cast<const uint8_t*>(after + 2 * 8 * kRingStages);
    R.stage_u32 = smem_u32(R.stage_buf);
    R.nst = a.ring_depth + 1;
    R.stage = 0;
    R.parity = 0;
    R.tiles = 0;
    if (threadIdx.x == 0) {
      for (int k = 0; k < kRingStages; ++k) {
        mbar_init(R.full0 + 8u * k, kWarpsPerBlock);   // every warp arrives with the bytes of the chunks it requested
        mbar_init(R.empty0 + 8u * k, kWarpsPerBlock);  // every warp arrives when it is done reading the tile
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  } else if (a.staged) {  // per-warp stage buffers + mbarriers for the SELL slices (values + columns of a slice are contiguous)
    char* smem = reinterpret_cast<char*>(dyn_smem);
    S.base = smem + warp * 2 * kStageBytes;
    S.bar[0] = smem_u32(smem + kWarpsPerBlock * 2 * kStageBytes + warp * 16);
    S.bar[1] = S.bar[0] + 8;
    S.count = 0;
    if (lane == 0) {
      mbar_init(S.bar[0], 1);
      mbar_init(S.bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  int nstamp = 0;
  stamp(a, nstamp);
  int cur = 0;  // which of the two tagged ghost buffers the current search direction uses (MULTI)

  // owned value of the search direction: plain store, and straight into the ghost slot of every neighbour rank that needs it
  auto publish_p = [&](int64_t row, double v, bool slice_sends) {
    vp[row] = v;
    if constexpr (MULTI) {
      if (slice_sends) {
        for (int e = __ldg(a.send_of_row + row); e >= 0;) {
          const SendEnt se = a.send_ents[e];
          st_tag<true>(se.t[cur], v, vtag);
          e = se.next;
        }
      }
    }
  };
  auto sends = [&](int64_t row) { return MULTI ? __ldg(a.slice_send + row / kSlice) != 0 : false; };
  // one SELL row (direct loads) times p
  auto sell_times_p = [&](int64_t s, int64_t row) {
    const int64_t beg = __ldg(a.slice_ptr + s);
    const int width = (int)((__ldg(a.slice_ptr + s + 1) - beg) / kSlice);
    return row_times_p<MULTI>(
        a, width, [&](int k) { return __ldg(a.cols + beg + (int64_t)k * kSlice + lane); },
        [&](int k) { return __ldg(a.A + beg + (int64_t)k * kSlice + lane); }, vp, a.tb[cur], vtag, &sh.fail);
  };

  // ---- K2 + initial residual: b = B v_ (+ dt stim) ; r = b - A x0 ; z = D^-1 r ; p = z -----------------------------
  double acc3[3] = {0.0, 0.0, 0.0};  // r.z, norm^2 of r, norm^2 of b (chosen norm)
  auto rhs_finish = [&](int64_t row, double bi, double ax0, double di, double sv) {
    if (a.has_stim) bi = fma(a.dt, sv, bi);
    const double ri = x0_prev ? bi - ax0 : bi;
    const double zi = di * ri;
    __stcg(vx + row, x0_prev ? __ldg(a.v_prev + row) : 0.0);
    __stcg(vr + row, ri);
    publish_p(row, zi, sends(row));
    acc3[0] = fma(ri, zi, acc3[0]);
    acc3[1] += norm_term(a.norm_type, ri, zi);
    acc3[2] += norm_term(a.norm_type, bi, di * bi);
  };
  if constexpr (MODE == kModeRing) {
    // rows outside the dictionary first (side table, thread per row, entries in the SELL order)
    for (int64_t i = __ldg(a.nd_cta_ptr + blockIdx.x) + threadIdx.x; i < __ldg(a.nd_cta_ptr + blockIdx.x + 1); i += kPdeThreads) {
      const int64_t row = __ldg(a.nd_rows + i);
      const int w = __ldg(a.nd_w + i);
      double bi = 0.0, ax0 = 0.0;
      for (int k = 0; k < w; ++k) {
        const double g = __ldg(a.v_prev + __ldg(a.nd_cols + i * kChunk + k));
        bi = fma(__ldg(a.nd_B + i * kChunk + k), g, bi);
        if (x0_prev) ax0 = fma(__ldg(a.nd_A + i * kChunk + k), g, ax0);
      }
      rhs_finish(row, bi, ax0, __ldg(a.dinv + row), a.has_stim ? __ldg(a.stim_vec + row) : 0.0);
    }
    PatRegs P;
    P.id = -1;
    const unsigned mask = (1u << R.cap_log2) - 1u;
    ring_sweep(a, R, a.v_prev, a.n_owned, row_b, row_e, [&](int64_t row, unsigned tok, int pid, int stim_flag) {
      if (row < row_e && pid != 255) {
        if (pid != P.id) pat_load(P, pid, D.sB);
        const double bi = ring_row(P, sEnc, R.buf_u32, row, mask, tok);
        double ax0 = 0.0;
        if (x0_prev) {
          const uint2* e = reinterpret_cast<const uint2*>(sEnc) + pid * kChunk;
#pragma unroll
          for (int u = 0; u < kChunk; ++u) ax0 = fma(D.sA[pid * kChunk + u], ring_at(R.buf_u32, e[u], (unsigned)row, mask, tok), ax0);
        }
        // the Jacobi diagonal of a dictionary row is a property of its stencil; stim_vec is read only where it is non-zero
        rhs_finish(row, bi, ax0, sDinv[pid], (a.has_stim && stim_flag) ? __ldg(a.stim_vec + row) : 0.0);
      }
    });
  } else {
    const bool staged_rhs = !DICT && a.staged && !x0_prev;
    if (staged_rhs && lane == 0) {
      if (s_first < s_end) S.issue(a, a.B, s_first, S.count);
      if (s_first + W < s_end) S.issue(a, a.B, s_first + W, S.count + 1);
    }
    for (int64_t s = s_first; s < s_end; s += W) {
      RowRef r;
      r.row = s * kSlice + lane;
      r.lane = lane;
      r.slot = 0;
      const bool on = r.row < a.n_owned;
      double bi = 0.0, ax0 = 0.0;
      if constexpr (DICT) {
        if (on) {
          r.beg = __ldg(a.slice_ptr + s);
          r.width = (int)((__ldg(a.slice_ptr + s + 1) - r.beg) / kSlice);
          dict_rhs_row<false>(a, D, r, x0_prev, bi, ax0);
        }
      } else if (staged_rhs) {
        bi = staged_row<false, false>(a, S, a.B, s, s + 2 * W < s_end ? s + 2 * W : -1, lane, a.v_prev, 0, &sh.fail);
      } else {
        r.beg = __ldg(a.slice_ptr + s);
        r.width = (int)((__ldg(a.slice_ptr + s + 1) - r.beg) / kSlice);
        rhs_row<false>(a, r, x0_prev, bi, ax0);
      }
      if (on) rhs_finish(r.row, bi, ax0, __ldg(a.dinv + r.row), a.has_stim ? __ldg(a.stim_vec + r.row) : 0.0);
    }
  }
  stamp(a, nstamp);
  post<3, true>(acc3, a, gen, sh);   // also the barrier that makes p_0 visible
  wait<3, true>(acc3, a, gen++, sh);
  stamp(a, nstamp);
  double rz = acc3[0];
  double rnorm = sqrt(fabs(acc3[1]));
  const double ttol = fmax(a.rtol * sqrt(fabs(acc3[2])), a.atol);
  int its = 0;
  int reason = classify(rnorm, ttol, a.atol, 0, a.max_it, false);
  if (sh.fail) reason = MONO_KSP_DIVERGED_NAN;
  while (reason == 0) {
    // ---- K4a: q = A p ; p.q -------------------------------------------------------------------------------------
    double pq[1] = {0.0};
    if constexpr (MODE == kModeRing) {
      for (int64_t i = __ldg(a.nd_cta_ptr + blockIdx.x) + threadIdx.x; i < __ldg(a.nd_cta_ptr + blockIdx.x + 1); i += kPdeThreads) {
        const int64_t row = __ldg(a.nd_rows + i);
        const double qi = row_times_p<MULTI>(
            a, __ldg(a.nd_w + i), [&](int k) { return __ldg(a.nd_cols + i * kChunk + k); },
            [&](int k) { return __ldg(a.nd_A + i * kChunk + k); }, vp, a.tb[cur], vtag, &sh.fail);
        __stcg(vq + row, qi);
        pq[0] = fma(__ldcg(vp + row), qi, pq[0]);
      }
      const unsigned mask = (1u << R.cap_log2) - 1u;
      PatRegs PA;  // (scoped to this phase: registers the vector phases need for their loads in flight)
      PA.id = -1;
      ring_sweep(a, R, vp, a.n_owned, row_b, row_e, [&](int64_t row, unsigned tok, int pid, int) {
        if (row < row_e && pid != 255) {
          if (pid != PA.id) pat_load(PA, pid, D.sA);
          const double qi = (a.dbg & 2) ? 1.0 : ring_row(PA, sEnc, R.buf_u32, row, mask, tok);
          const double pi = R.buf[(a.ring_cl0 << R.cap_log2) + ((unsigned)row & mask)];  // the diagonal's cluster holds p[row]
          if (!(a.dbg & 4)) __stcg(vq + row, qi);
          pq[0] = fma(pi, qi, pq[0]);
        }
      });
    } else if constexpr (MODE == kModeDictL1) {
      // two slices of the warp per trip: the pattern bytes, then up to 2 x 16 gathers, are in flight together
      for (int64_t s = s_first; s < s_end; s += 2 * W) {
        const int64_t row0 = s * kSlice + lane, row1 = (s + W) * kSlice + lane;
        const bool on0 = row0 < a.n_owned, on1 = s + W < s_end && row1 < a.n_owned;
        const int p0 = on0 ? (int)__ldg(a.pat + row0) : 255, p1 = on1 ? (int)__ldg(a.pat + row1) : 255;
        const double pi0 = on0 ? __ldcg(vp + row0) : 0.0, pi1 = on1 ? __ldcg(vp + row1) : 0.0;
        double q0 = 0.0, q1 = 0.0;
        if (p0 != 255 && p1 != 255) {
          const int b0 = p0 * kChunk, b1 = p1 * kChunk, w0 = D.sW[p0], w1 = D.sW[p1];
          double g0[kChunk], g1[kChunk];
#pragma unroll
          for (int u = 0; u < kChunk; ++u) g0[u] = u < w0 ? ld_coherent(vp + row0 + D.sOff[b0 + u]) : 0.0;
#pragma unroll
          for (int u = 0; u < kChunk; ++u) g1[u] = u < w1 ? ld_coherent(vp + row1 + D.sOff[b1 + u]) : 0.0;
#pragma unroll
          for (int u = 0; u < kChunk; ++u) q0 = fma(u < w0 ? D.sA[b0 + u] : 0.0, g0[u], q0);
#pragma unroll
          for (int u = 0; u < kChunk; ++u) q1 = fma(u < w1 ? D.sA[b1 + u] : 0.0, g1[u], q1);
        } else {
          auto one = [&](bool on, int pid, int64_t ss, int64_t row) {
            if (!on) return 0.0;
            if (pid == 255) return sell_times_p(ss, row);
            const int b = pid * kChunk;
            return row_times_p<false>(
                a, D.sW[pid], [&](int k) { return (int32_t)row + D.sOff[b + k]; }, [&](int k) { return D.sA[b + k]; }, vp, nullptr, 0,
                &sh.fail);
          };
          q0 = one(on0, p0, s, row0);
          q1 = one(on1, p1, s + W, row1);
        }
        if (on0) {
          __stcg(vq + row0, q0);
          pq[0] = fma(pi0, q0, pq[0]);
        }
        if (on1) {
          __stcg(vq + row1, q1);
          pq[0] = fma(pi1, q1, pq[0]);
        }
      }
    } else if (a.staged) {
      if (lane == 0) {
        if (s_first < s_end) S.issue(a, a.A, s_first, S.count);
        if (s_first + W < s_end) S.issue(a, a.A, s_first + W, S.count + 1);
      }
      for (int64_t s = s_first; s < s_end; s += W) {
        const int64_t row = s * kSlice + lane;
        const bool on = row < a.n_owned;
        const double pi = on ? __ldcg(vp + row) : 0.0;  // requested before the gathers
        const double qi = staged_row_p<MULTI>(a, S, a.A, s, s + 2 * W < s_end ? s + 2 * W : -1, lane, vp, a.tb[cur], vtag, &sh.fail);
        if (on) {
          __stcg(vq + row, qi);
          pq[0] = fma(pi, qi, pq[0]);
        }
      }
    } else {
      for (int64_t s = s_first; s < s_end; s += W) {
        const int64_t row = s * kSlice + lane;
        if (row < a.n_owned) {
          const double pi = __ldcg(vp + row);
          const double qi = sell_times_p(s, row);
          __stcg(vq + row, qi);
          pq[0] = fma(pi, qi, pq[0]);
        }
      }
    }
    stamp(a, nstamp);
    post<1>(pq, a, gen, sh);
    wait<1>(pq, a, gen++, sh);
    stamp(a, nstamp);
    const double alpha = rz / pq[0];
    // ---- K4b: x += alpha p ; r -= alpha q ; z = D^-1 r ; r.z and the residual norm: four rows per trip ---------------
    double acc2[2] = {0.0, 0.0};
    for (int64_t s = s_first; s < s_end; s += 4 * W) {
      double q[4], r[4], d[4], pp[4], x[4];
      bool on[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t row = (s + j * W) * kSlice + lane;
        on[j] = s + j * W < s_end && row < a.n_owned;
        if (on[j]) {
          q[j] = __ldcg(vq + row);
          r[j] = __ldcg(vr + row);
          d[j] = __ldcg(a.dinv + row);
          pp[j] = __ldcg(vp + row);
          x[j] = __ldcg(vx + row);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (on[j]) {
          const int64_t row = (s + j * W) * kSlice + lane;
          const double ri = fma(-alpha, q[j], r[j]);
          const double zi = d[j] * ri;
          __stcg(vx + row, fma(alpha, pp[j], x[j]));
          __stcg(vr + row, ri);
          acc2[0] = fma(ri, zi, acc2[0]);
          acc2[1] += norm_term(a.norm_type, ri, zi);
        }
      }
    }
    stamp(a, nstamp);
    post<2>(acc2, a, gen, sh);
    wait<2>(acc2, a, gen++, sh);
    stamp(a, nstamp);
    ++its;
    rnorm = sqrt(fabs(acc2[1]));
    const double beta = acc2[0] / rz;
    rz = acc2[0];
    reason = classify(rnorm, ttol, a.atol, its, a.max_it, !(pq[0] == pq[0]) || pq[0] == 0.0);
    if (sh.fail) reason = MONO_KSP_DIVERGED_NAN;
    if (reason != 0) break;
    // ---- p = z + beta p, then the barrier that makes it visible --------------------------------------------------
    ++vtag;
    cur ^= 1;
    for (int64_t s = s_first; s < s_end; s += 4 * W) {
      double pp[4], d[4], r[4];
      bool on[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t row = (s + j * W) * kSlice + lane;
        on[j] = s + j * W < s_end && row < a.n_owned;
        if (on[j]) {
          pp[j] = __ldcg(vp + row);
          d[j] = __ldcg(a.dinv + row);
          r[j] = __ldcg(vr + row);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (on[j]) {
          const int64_t sj = s + j * W;
          publish_p(sj * kSlice + lane, fma(beta, pp[j], d[j] * r[j]), MULTI ? __ldg(a.slice_send + sj) != 0 : false);
        }
      }
    }
    stamp(a, nstamp);
    double one[1] = {0.0};
    post<1, true>(one, a, gen, sh);
    wait<1, true>(one, a, gen++, sh);
    stamp(a, nstamp);
  }
  // ---- end of the solve: x is in place; ghost refresh of x (state.x.scatter_forward(), base_model.py:242) -----------
  if constexpr (MULTI) {
    const u64 xtag = gen;
    for (int64_t s = s_first; s < s_end; s += W) {
      const int64_t row = s * kSlice + lane;
      if (row < a.n_owned && __ldg(a.slice_send + s) != 0) {
        int e = __ldg(a.send_of_row + row);
        if (e >= 0) {
          const double xi = __ldcg(vx + row);
          while (e >= 0) {
            const SendEnt se = a.send_ents[e];
            st_tag<true>(se.xg, xi, xtag);
            e = se.next;
          }
        }
      }
    }
    const int64_t n_ghost = a.n_local - a.n_owned;
    for (int64_t g = (int64_t)blockIdx.x * kPdeThreads + threadIdx.x; g < n_ghost; g += (int64_t)a.n_workers * kPdeThreads)
      a.x[a.n_owned + g] = wait_tag<true>(a.xg + g, xtag, &sh.fail, a.spin_ns);
  }
  if (sh.fail && threadIdx.x == 0) a.res->error = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) a.gen_state[1] = vtag;  // every CTA read it at its start
}

// ---- KSPPIPECG: one reduction per iteration, overlapped with the SpMV -----------------------------------------
// Ghysels & Vanroose, "Hiding global synchronization latency in the preconditioned Conjugate Gradient
// algorithm" (Alg. 3), as in PETSc's KSPPIPECG:
//   gamma = (r,u) ; delta = (w,u) ; m = M^-1 w ; n = A m ; beta = gamma/gamma_old ;
//   alpha = gamma / (delta - beta*gamma/alpha_old) ; z = n + beta z ; q = m + beta q ; s = w + beta s ;
//   p = u + beta p ; x += alpha p ; r -= alpha s ; u -= alpha q ; w -= alpha z
// The partial sums of gamma, delta are posted BEFORE m and n are computed and the totals are awaited after.
//
// Preconditioner M^-1 = p(D^-1 A) D^-1: cheb_k steps of the Chebyshev iteration for D^-1 A y = D^-1 w on
// [b/kappa, b] started from zero (k = 1: plain Jacobi up to a scale; b = Gershgorin bound of D^-1 A, which keeps
// M^-1 positive definite whatever kappa is).  Every step is one more SpMV whose operand is exchanged through its own
// tagged buffer - no reduction - so a degree-k preconditioner trades k-1 cheap dataflow-synchronised SpMVs for
// ~k times fewer grid-wide reductions, which is what a small mesh is bound by (58 176 rows: 7 -> 3 iterations).
// Buffers: set s, step j -> tb[s*k + j]; applications alternate between sets 0 and 1 (the reduction of every
// iteration is the all-to-all dependency that makes the reuse safe), set 2 serves M^-1 b at start-up.
template <bool RESIDENT, bool MATSMEM, bool MULTI, bool CHEB>
__global__ void __launch_bounds__(kPdeThreads, 1) pde_pipecg_kernel(const PdeArgs a) {
  extern __shared__ double dyn_smem[];
  __shared__ Scratch sh;
  if (threadIdx.x == 0) sh.fail = 0;
  __syncthreads();
  u64 gen = a.gen_state[0];
  u64 vtag = a.gen_state[1] + 1;  // tag of the vector published last (every CTA publishes the same sequence)
  if (blockIdx.x == a.n_workers) {
    pipecg_reducer<MULTI>(a, sh, gen);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)(threadIdx.x >> 5) * a.n_workers + blockIdx.x;
  const int64_t warp_stride = (int64_t)a.n_workers * kWarpsPerBlock;
  const bool x0_prev = a.x0_mode == MONO_X0_PREVIOUS;
  const int K = CHEB ? a.cheb_k : 1;  // compile-time 1 for plain Jacobi: the polynomial code drops out of that build
  constexpr bool cheb = CHEB;
  constexpr int KIND = CHEB ? 2 : 1;
  const VecStore<RESIDENT, KIND> V = make_store<RESIDENT, KIND>(a, dyn_smem);
  double* sa = dyn_smem + (size_t)kNvec[KIND] * V.cap;
  int32_t* sc = reinterpret_cast<int32_t*>(sa + kChunk * kPdeThreads);
  const int width_cached = stage_matrix<MATSMEM>(a, sa, sc, warp_global, lane);
  const MatA<MATSMEM, CHEB> Aop{a, sa, sc};
  int nstamp = 0;
  stamp(a, nstamp);

  // first Chebyshev iterate of an application to the vector whose Jacobi-scaled own value is gi; published in (set, 0)
  auto start_app = [&](const RowRef& r, int set, double gi, u64 tag) {
    if (cheb) {
      gi *= a.cheb_inv_theta;
      V.st(VY, r, gi);
      V.st(VE, r, gi);
    }
    publish<MULTI>(a, set * K, r.row, gi, tag);
  };
  // steps 1 .. K-1 of the application started in `set` (operand D^-1 * vector SRC); leaves y_K in VY and in (set, K-1)
  auto cheb_steps = [&](int set, int SRC) {
    for (int j = 1; j < K; ++j) {
      const double c1 = a.cheb_c1[j], c2 = a.cheb_c2[j];
      OWN_ROWS_BEGIN
        const double ti = Aop.template apply<MULTI>(r, a.tb[set * K + j - 1], vtag, &sh.fail);
        if (r.row < a.n_owned) {
          const double di = V.ld(VD, r);
          const double dj = fma(c1, V.ld(VE, r), c2 * (di * (V.ld(SRC, r) - ti)));
          const double yi = V.ld(VY, r) + dj;
          V.st(VE, r, dj);
          V.st(VY, r, yi);
          publish<MULTI>(a, set * K + j, r.row, yi, vtag + 1);
        }
      OWN_ROWS_END
      ++vtag;
    }
  };

  // ---- P0: b = B v_ (+ stimulus) ; r = b - A x0 ; first iterate of u = M^-1 r -> set 1 -------------------------
  const bool need_mb = cheb && a.norm_type != MONO_NORM_UNPRECONDITIONED;  // |b| is measured through M^-1
  const bool app_b = need_mb && x0_prev;                                    // r0 != b: M^-1 b needs its own application
  double acc[4] = {0.0, 0.0, 0.0, 0.0};  // gamma, delta, norm^2 of r, norm^2 of b (chosen norm)
  OWN_ROWS_BEGIN
    RowRef g = r;
    if constexpr (MATSMEM) {
      g.beg = __ldg(a.slice_ptr + s__);
    }
    double bi, ax0;
    rhs_row(a, g, x0_prev, bi, ax0);
    if (r.row < a.n_owned) {
      const double di = __ldg(a.dinv + r.row);
      if constexpr (RESIDENT) V.st(VD, r, di);
      const double ri = x0_prev ? bi - ax0 : bi;
      V.st(VX, r, x0_prev ? __ldg(a.v_prev + r.row) : 0.0);
      V.st(VR, r, ri);
      if (!need_mb) acc[3] += norm_term(a.norm_type, bi, di * bi);
      if (app_b) {
        V.st(VS, r, bi);
        start_app(r, 2, di * bi, vtag);
      } else {
        if (!cheb) V.st(VU, r, di * ri);
        start_app(r, 1, di * ri, vtag);
      }
    }
  OWN_ROWS_END
  stage_wait<MATSMEM>();
  if (app_b) {  // M^-1 b (set 2), then the first iterate of u = M^-1 r0 (set 1)
    cheb_steps(2, VS);
    ++vtag;
    OWN_ROWS_BEGIN
      if (r.row < a.n_owned) {
        acc[3] += norm_term(a.norm_type, V.ld(VS, r), V.ld(VY, r));
        start_app(r, 1, V.ld(VD, r) * V.ld(VR, r), vtag);
      }
    OWN_ROWS_END
  }
  cheb_steps(1, VR);
  stamp(a, nstamp);

  // ---- P1: w = A u ; first iterate of m = M^-1 w -> set 0 ; gamma, delta, norms ------------------------------------
  OWN_ROWS_BEGIN
    const double wi = Aop.template apply<MULTI>(r, a.tb[1 * K + K - 1], vtag, &sh.fail);
    if (r.row < a.n_owned) {
      const double ri = V.ld(VR, r);
      double ui;
      if (cheb) {
        ui = V.ld(VY, r);
        V.st(VU, r, ui);
      } else {
        ui = V.ld(VU, r);
      }
      V.st(VW, r, wi);
      start_app(r, 0, V.ld(VD, r) * wi, vtag + 1);
      acc[0] = fma(ri, ui, acc[0]);
      acc[1] = fma(wi, ui, acc[1]);
      acc[2] += norm_term(a.norm_type, ri, ui);
      if (need_mb && !app_b) acc[3] += norm_term(a.norm_type, ri, ui);  // x0 = 0: b = r0, M^-1 b = u
    }
  OWN_ROWS_END
  ++vtag;
  stamp(a, nstamp);

  int its = 0, reason = 0;
  double rnorm = 0.0, gamma_old = 1.0, alpha_old = 1.0, ttol = 0.0;
  int cur = 0;  // set that holds the current application (m)
  while (true) {
    if (its == 0)
      post<4>(acc, a, gen, sh);
    else
      post<3>(reinterpret_cast<const double(&)[3]>(acc), a, gen, sh);
    stamp(a, nstamp);
    // m = M^-1 w (remaining Chebyshev steps) and n = A m while the reduction is in flight
    cheb_steps(cur, VW);
    OWN_ROWS_BEGIN
      const double ni = Aop.template apply<MULTI>(r, a.tb[cur * K + K - 1], vtag, &sh.fail);
      if (r.row < a.n_owned) V.st(VN, r, ni);
    OWN_ROWS_END
    stamp(a, nstamp);
    if (its == 0) {
      wait<4>(acc, a, gen++, sh);
      ttol = fmax(a.rtol * sqrt(fabs(acc[3])), a.atol);
    } else {
      wait<3>(reinterpret_cast<double(&)[3]>(acc), a, gen++, sh);
    }
    stamp(a, nstamp);
    const double gamma = acc[0], delta = acc[1];
    rnorm = sqrt(fabs(acc[2]));
    reason = classify(rnorm, ttol, a.atol, its, a.max_it, !(gamma == gamma) || !(delta == delta));
    if (sh.fail) reason = MONO_KSP_DIVERGED_NAN;
    if (reason != 0) break;
    const bool first = its == 0;
    const double beta = first ? 0.0 : gamma / gamma_old;
    const double alpha = first ? gamma / delta : gamma / (delta - beta * gamma / alpha_old);
    acc[0] = acc[1] = acc[2] = 0.0;
    OWN_ROWS_BEGIN
      if (r.row < a.n_owned) {
        const double di = V.ld(VD, r);
        double xi = V.ld(VX, r), ri = V.ld(VR, r), ui = V.ld(VU, r), wi = V.ld(VW, r);
        const double mi = cheb ? V.ld(VY, r) : di * wi;
        double zi = V.ld(VN, r), qi = mi, si = wi, pi = ui;
        if (!first) {
          zi = fma(beta, V.ld(VZ, r), zi);
          qi = fma(beta, V.ld(VQ, r), qi);
          si = fma(beta, V.ld(VS, r), si);
          pi = fma(beta, V.ld(VP, r), pi);
        }
        xi = fma(alpha, pi, xi);
        ri = fma(-alpha, si, ri);
        ui = fma(-alpha, qi, ui);
        wi = fma(-alpha, zi, wi);
        V.st(VZ, r, zi);
        V.st(VQ, r, qi);
        V.st(VS, r, si);
        V.st(VP, r, pi);
        V.st(VX, r, xi);
        V.st(VR, r, ri);
        V.st(VU, r, ui);
        V.st(VW, r, wi);
        start_app(r, cur ^ 1, di * wi, vtag + 1);
        acc[0] = fma(ri, ui, acc[0]);
        acc[1] = fma(wi, ui, acc[1]);
        acc[2] += norm_term(a.norm_type, ri, ui);
      }
    OWN_ROWS_END
    ++vtag;
    cur ^= 1;
    gamma_old = gamma;
    alpha_old = alpha;
    ++its;
  }
  FINISH_SOLVE(gen)
}

// measurement: `n` back-to-back reductions (cost of one grid-wide reduction of 3 scalars)
__global__ void __launch_bounds__(kPdeThreads, 1) sync_bench_kernel(const PdeArgs a, int n, double* out) {
  __shared__ Scratch sh;
  if (threadIdx.x == 0) sh.fail = 0;
  __syncthreads();
  u64 gen = a.gen_state[0];
  double v[3] = {1.0, 2.0, 3.0};
  for (int i = 0; i < n; ++i) {
    if (blockIdx.x == a.n_workers) {
      reduce_publish<3, false>(v, a, gen++, sh);
    } else {
      post<3>(v, a, gen, sh);
      wait<3>(v, a, gen++, sh);
    }
    v[0] = v[0] * 1e-3 + 1.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = v[0] + sh.fail;
  if (blockIdx.x == a.n_workers && threadIdx.x == 0) a.gen_state[0] = gen;
}

// ---- K3: A = C_m*Mass + dt*theta*K ; B = C_m*Mass - dt*(1-theta)*K ; Jacobi diagonal ------------------
__global__ void build_ab_kernel(int64_t nnz, const double* __restrict__ mass, const double* __restrict__ stiff,
                                double* __restrict__ A, double* __restrict__ B, double cm, double a_k, double b_k) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    const double m = cm * mass[e], k = stiff[e];
    A[e] = fma(a_k, k, m);
    B[e] = fma(-b_k, k, m);
  }
}

// also: *gersh = max_i sum_j |a_ij| / a_ii, the Gershgorin bound of the spectrum of D^-1 A (upper end of the
// Chebyshev interval; positive doubles order like their bit patterns, so an integer atomicMax does it)
// dictionary values = the A / B entries of each pattern's representative row (same bits as the SELL arrays hold)
__global__ void dict_gather_kernel(int n, const int64_t* __restrict__ src, const double* __restrict__ A, const double* __restrict__ B,
                                   double* __restrict__ dA, double* __restrict__ dB) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int64_t e = src[i];
    dA[i] = e >= 0 ? A[e] : 0.0;
    dB[i] = e >= 0 ? B[e] : 0.0;
  }
}

__global__ void dict_dinv_kernel(int n, const int64_t* __restrict__ rep, const double* __restrict__ dinv, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dinv[rep[i]];
}

__global__ void jacobi_kernel(int64_t n_owned, int64_t n_slices, const int64_t* __restrict__ slice_ptr,
                              const int32_t* __restrict__ cols, const double* __restrict__ A,
                              double* __restrict__ dinv, int pc_type, unsigned long long* __restrict__ gersh) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_owned; row += stride) {
    if (pc_type == MONO_PC_NONE) {
      dinv[row] = 1.0;
      continue;
    }
    const int64_t s = row / kSlice;
    const int lane = (int)(row % kSlice);
    const int64_t beg = slice_ptr[s];
    const int width = (int)((slice_ptr[s + 1] - beg) / kSlice);
    double d = 0.0, sabs = 0.0;
    for (int k = 0; k < width; ++k) {
      const int64_t e = beg + (int64_t)k * kSlice + lane;
      if (cols[e] == row) d += A[e];
      sabs += fabs(A[e]);
    }
    dinv[row] = 1.0 / d;
    if (gersh != nullptr && d > 0.0) atomicMax(gersh, (unsigned long long)__double_as_longlong(sabs / d));
  }
}

// stim_vec[idx[e]] (+)= scale * val[e]   (scale == 0: clear the entries of a stimulus that went inactive)
// (slice_flag: per SELL slice, set where an active stimulus touches a row - the ring kernel reads stim_vec only there)
__global__ void stim_scatter_kernel(int64_t nnz, const int32_t* __restrict__ idx, const double* __restrict__ val,
                                    double scale, double* __restrict__ stim_vec, uint8_t* __restrict__ slice_flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    if (scale == 0.0) {
      stim_vec[idx[e]] = 0.0;
    } else {
      atomicAdd(stim_vec + idx[e], scale * val[e]);
      slice_flag[idx[e] / kSlice] = 1;
    }
  }
}

// order-preserving map double -> uint64 (so that atomicMin/atomicMax on integers order like the doubles)
__device__ __forceinline__ unsigned long long ordered_bits(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// Observers after a split step, one launch: point probes (+ their activation times), per-node activation map,
// min / max of v over the owned nodes.
__global__ void observers_kernel(int n_probes, const ProbeDev* __restrict__ probes, const double* __restrict__ x,
                                 double* __restrict__ vals, double* __restrict__ act, int act_enabled, double threshold, double t0,
                                 int64_t n_owned, double* __restrict__ actmap, double map_threshold, unsigned long long* __restrict__ minmax) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < n_probes) {
    const ProbeDev pr = probes[tid];
    double v = 0.0;
    for (int k = 0; k < pr.n; ++k) v = fma(pr.w[k], x[pr.node[k]], v);
    vals[tid] = v;
    if (act_enabled && act[tid] < 0.0 && v > threshold) act[tid] = t0;
  }
  if (actmap == nullptr && minmax == nullptr) return;
  double lo = INFINITY, hi = -INFINITY;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n_owned; i += stride) {
    const double v = x[i];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
    if (actmap != nullptr && v > map_threshold && actmap[i] < 0.0) actmap[i] = t0;
  }
  if (minmax != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
      atomicMin(minmax, ordered_bits(lo));
      atomicMax(minmax + 1, ordered_bits(hi));
    }
  }
}

__global__ void strided_pack_kernel(int64_t count, int64_t offset, int64_t stride, const double* __restrict__ x, double* __restrict__ out) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += step) out[k] = x[offset + k * stride];
}

}  // namespace

// Stencil dictionary of the matrices (default; MONO_PDE_DICT=0 disables), see dict_apply.  Keeps the kMaxPat most frequent
// stencils of at most kChunk entries; the dictionary is used when it covers at least half of the rows.
static int pde_build_dictionary(mono_ctx* c, const int64_t* indptr, const int32_t* indices, const double* mass, const double* stiff,
                                const std::vector<int64_t>& sp) {
  const int64_t n = c->n_owned;
  std::vector<uint8_t> pat((size_t)n);
  int32_t npat = 0;
  int64_t rep[kMaxPat], cnt[kMaxPat];
  if (mono_csr_row_patterns(n, indptr, indices, mass, stiff, kMaxPat, pat.data(), &npat, rep, cnt) != MONO_OK)
    return mono_fail(c, MONO_E_INVALID, std::string("stencil dictionary: ") + mono_last_error(nullptr));
  // drop patterns wider than one batch of sell_row and renumber
  int remap[256];
  for (int i = 0; i < 256; ++i) remap[i] = 255;
  std::vector<int64_t> krep;
  int64_t covered = 0;
  for (int p = 0; p < npat; ++p) {
    if (indptr[rep[p] + 1] - indptr[rep[p]] > kChunk) continue;
    bool ghost = false;  // a stencil that reaches into the ghost block (row next to a partition cut) is not a dictionary stencil
    for (int64_t k = indptr[rep[p]]; k < indptr[rep[p] + 1]; ++k) ghost |= indices[k] >= n;
    if (ghost) continue;
    remap[p] = (int)krep.size();
    krep.push_back(rep[p]);
    covered += cnt[p];
  }
  c->dict_cover = (double)covered / (double)n;
  if (krep.empty() || c->dict_cover < 0.5) {
    c->n_pat = 0;
    return MONO_OK;
  }
  // rows with a ghost column stay on the SELL path: a dictionary row gathers only from the plain owned vector
  for (int64_t r = 0; r < n; ++r) {
    uint8_t q = (uint8_t)remap[pat[(size_t)r]];
    if (q != 255)
      for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k)
        if (indices[k] >= n) {
          q = 255;
          --covered;
          break;
        }
    pat[(size_t)r] = q;
  }
  c->dict_cover = (double)covered / (double)n;
  if (c->dict_cover < 0.5) {
    c->n_pat = 0;
    return MONO_OK;
  }
  const int np = (int)krep.size();
  std::vector<int32_t> off((size_t)np * kChunk, 0), w((size_t)np, 0);
  std::vector<int64_t> src((size_t)np * kChunk, -1);
  for (int p = 0; p < np; ++p) {
    const int64_t r = krep[(size_t)p];
    const int wd = (int)(indptr[r + 1] - indptr[r]);
    w[(size_t)p] = wd;
    for (int k = 0; k < wd; ++k) {
      off[(size_t)p * kChunk + k] = indices[indptr[r] + k] - (int32_t)r;
      src[(size_t)p * kChunk + k] = sp[(size_t)(r / kSlice)] + (int64_t)k * kSlice + (r % kSlice);
    }
  }
  // ---- clusters of the column offsets (the planes of a structured mesh) for the shared-memory rings -----------------
  // one ring of 2^cap_log2 elements per cluster has to hold its span plus the tiles in flight; pick the clustering (gap
  // threshold) that needs the least shared memory
  std::vector<int32_t> cl((size_t)np * kChunk, 0);
  c->ring_ok = false;
  {
    std::vector<int64_t> offs;
    for (int p = 0; p < np; ++p)
      for (int k = 0; k < w[(size_t)p]; ++k) offs.push_back(off[(size_t)p * kChunk + k]);
    std::sort(offs.begin(), offs.end());
    offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
    size_t best_bytes = 0;
    for (int64_t gap : {64ll, 256ll, 1024ll, 4096ll, 16384ll, 65536ll}) {
      std::vector<std::pair<int64_t, int64_t>> cls;
      for (int64_t o : offs) {
        if (cls.empty() || o - cls.back().second > gap)
          cls.emplace_back(o, o);
        else
          cls.back().second = o;
      }
      if ((int)cls.size() > kMaxClusters) continue;
      int64_t span = 0;
      for (auto& q : cls) span = std::max(span, q.second - q.first + 1);
      int lg = 9;
      while ((1ll << lg) < span + (int64_t)(kRingMinDepth + 1) * kRingTile + kRingChunk) ++lg;
      auto fits = [&](int l) { return cls.size() * ((size_t)8 << l) + kRingTableSmem + (size_t)kRingStages * (kRingStageBytes + 16) + 256 <= 218 * 1024; };
      while (((1ll << lg) - span - kRingChunk) / kRingTile - 1 < 3 && fits(lg + 1)) ++lg;  // three tiles ahead if they fit
      const size_t bytes = cls.size() * ((size_t)8 << lg);
      if (bytes + kRingTableSmem + (size_t)kRingStages * (kRingStageBytes + 16) + 256 > 218 * 1024) continue;
      if (!c->ring_ok || bytes < best_bytes) {
        c->ring_ok = true;
        best_bytes = bytes;
        c->ring_ncl = (int)cls.size();
        c->ring_cap_log2 = lg;
        // the power of two usually leaves room for more tiles in flight than the minimum
        c->ring_depth = (int)std::min<int64_t>(kRingMaxDepth, ((1ll << lg) - span - kRingChunk) / kRingTile - 1);
        for (size_t q = 0; q < cls.size(); ++q) {
          c->ring_lo[q] = cls[q].first;
          c->ring_hi[q] = cls[q].second;
        }
      }
    }
    if (c->ring_ok) {
      int cl0 = 0;  // cluster of offset 0 (the diagonal): padding entries point there
      for (int q = 0; q < c->ring_ncl; ++q)
        if (c->ring_lo[q] <= 0 && 0 <= c->ring_hi[q]) cl0 = q;
      for (int p = 0; p < np; ++p)
        for (int k = 0; k < kChunk; ++k) {
          int q = cl0;
          if (k < w[(size_t)p])
            for (int j = 0; j < c->ring_ncl; ++j)
              if (c->ring_lo[j] <= off[(size_t)p * kChunk + k] && off[(size_t)p * kChunk + k] <= c->ring_hi[j]) q = j;
          cl[(size_t)p * kChunk + k] = q;
        }
    }
  }
  // ---- side table: the rows outside the dictionary, row-major (16 entries each) --------------------------------------
  c->nd_rows_host.clear();
  for (int64_t r = 0; r < n; ++r)
    if (pat[(size_t)r] == 255) c->nd_rows_host.push_back((int32_t)r);
  c->n_nd = (int64_t)c->nd_rows_host.size();
  bool side_ok = true;
  for (int32_t r : c->nd_rows_host)
    if (indptr[r + 1] - indptr[r] > kChunk) side_ok = false;
  if (!side_ok || c->n_nd > n / 4) c->ring_ok = false;  // (wide or many irregular rows: the SELL stream serves them better)
  if (c->ring_ok && c->n_nd > 0) {
    const size_t m = (size_t)c->n_nd;
    std::vector<int32_t> ncols(m * kChunk), nw(m);
    std::vector<int64_t> nsrc(m * kChunk, -1);
    for (size_t i = 0; i < m; ++i) {
      const int64_t r = c->nd_rows_host[i];
      const int wd = (int)(indptr[r + 1] - indptr[r]);
      nw[i] = wd;
      for (int k = 0; k < kChunk; ++k) {
        ncols[i * kChunk + k] = k < wd ? indices[indptr[r] + k] : (int32_t)r;
        if (k < wd) nsrc[i * kChunk + k] = sp[(size_t)(r / kSlice)] + (int64_t)k * kSlice + (r % kSlice);
      }
    }
    MONO_CUDA(c, cudaMalloc(&c->nd_rows_dev, m * sizeof(int32_t)));
    MONO_CUDA(c, cudaMalloc(&c->nd_w_dev, m * sizeof(int32_t)));
    MONO_CUDA(c, cudaMalloc(&c->nd_cols_dev, m * kChunk * sizeof(int32_t)));
    MONO_CUDA(c, cudaMalloc(&c->nd_src_dev, m * kChunk * sizeof(int64_t)));
    MONO_CUDA(c, cudaMalloc(&c->nd_A_dev, m * kChunk * sizeof(double)));
    MONO_CUDA(c, cudaMalloc(&c->nd_B_dev, m * kChunk * sizeof(double)));
    MONO_CUDA(c, cudaMemcpyAsync(c->nd_rows_dev, c->nd_rows_host.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaMemcpyAsync(c->nd_w_dev, nw.data(), m * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaMemcpyAsync(c->nd_cols_dev, ncols.data(), m * kChunk * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaMemcpyAsync(c->nd_src_dev, nsrc.data(), m * kChunk * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  MONO_CUDA(c, cudaMalloc(&c->pat_dev, (size_t)n + 2 * kRingTile));  // (padded: the ring kernel copies whole tiles)
  MONO_CUDA(c, cudaMemsetAsync(c->pat_dev, 0xff, (size_t)n + 2 * kRingTile, c->stream));
  MONO_CUDA(c, cudaMalloc(&c->dict_rep_dev, krep.size() * sizeof(int64_t)));
  MONO_CUDA(c, cudaMalloc(&c->dict_dinv_dev, krep.size() * sizeof(double)));
  MONO_CUDA(c, cudaMemcpyAsync(c->dict_rep_dev, krep.data(), krep.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMalloc(&c->dict_off_dev, off.size() * sizeof(int32_t)));
  MONO_CUDA(c, cudaMalloc(&c->dict_cl_dev, cl.size() * sizeof(int32_t)));
  MONO_CUDA(c, cudaMalloc(&c->dict_w_dev, w.size() * sizeof(int32_t)));
  MONO_CUDA(c, cudaMalloc(&c->dict_src_dev, src.size() * sizeof(int64_t)));
  MONO_CUDA(c, cudaMalloc(&c->dict_A_dev, off.size() * sizeof(double)));
  MONO_CUDA(c, cudaMalloc(&c->dict_B_dev, off.size() * sizeof(double)));
  MONO_CUDA(c, cudaMemcpyAsync(c->pat_dev, pat.data(), (size_t)n, cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->dict_off_dev, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->dict_cl_dev, cl.data(), cl.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->dict_w_dev, w.data(), w.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->dict_src_dev, src.data(), src.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_pat = np;
  return MONO_OK;
}

int pde_build_sell(mono_ctx* c, const int64_t* indptr, const int32_t* indices, const double* mass, const double* stiff) {
  const int64_t n = c->n_owned;
  const int64_t ns = (n + kSlice - 1) / kSlice;
  MONO_CHECK(c, indptr && (n == 0 || (indices && mass && stiff)), "CSR arrays are NULL");
  MONO_CHECK(c, indptr[0] == 0, "indptr[0] must be 0");
  for (int64_t r = 0; r < n; ++r) {
    MONO_CHECK(c, indptr[r + 1] >= indptr[r], "indptr is not non-decreasing");
    bool diag = false;
    for (int64_t k = indptr[r]; k < indptr[r + 1] && !diag; ++k) diag = indices[k] == r;
    MONO_CHECK(c, diag, "every owned row needs its diagonal entry (Jacobi scaling, mass matrix)");
  }
  std::vector<int64_t> sp(ns + 1, 0);
  for (int64_t s = 0; s < ns; ++s) {
    int64_t w = 0;
    for (int64_t r = s * kSlice; r < std::min(n, (s + 1) * kSlice); ++r) w = std::max(w, indptr[r + 1] - indptr[r]);
    sp[s + 1] = sp[s] + w * kSlice;
  }
  const int64_t tot = sp[ns];
  std::vector<int32_t> hc((size_t)tot);
  std::vector<double> hm((size_t)tot, 0.0), hk((size_t)tot, 0.0);
  for (int64_t s = 0; s < ns; ++s) {
    const int64_t w = (sp[s + 1] - sp[s]) / kSlice;
    for (int r = 0; r < kSlice; ++r) {
      const int64_t row = s * kSlice + r;
      const int64_t self = row < n ? row : n - 1;
      const int64_t rb = row < n ? indptr[row] : 0, re = row < n ? indptr[row + 1] : 0;
      for (int64_t k = 0; k < w; ++k) {
        const int64_t e = sp[s] + k * kSlice + r;
        if (k < re - rb) {
          const int32_t col = indices[rb + k];
          if (col < 0 || col >= c->n_local) return mono_fail(c, MONO_E_INVALID, "CSR column index out of range");
          hc[e] = col;
          hm[e] = mass[rb + k];
          hk[e] = stiff[rb + k];
        } else {
          hc[e] = (int32_t)self;
        }
      }
    }
  }
  c->n_slices = ns;
  c->sell_nnz = tot;
  c->max_width = 0;
  for (int64_t s = 0; s < ns; ++s) c->max_width = std::max<int>(c->max_width, (int)((sp[s + 1] - sp[s]) / kSlice));
  MONO_CUDA(c, cudaMalloc(&c->slice_ptr, (ns + 1) * sizeof(int64_t)));
  MONO_CUDA(c, cudaMalloc(&c->slice_stim_dev, (size_t)ns + 64));
  MONO_CUDA(c, cudaMemsetAsync(c->slice_stim_dev, 0, (size_t)ns + 64, c->stream));
  MONO_CUDA(c, cudaMalloc(&c->cols, std::max<int64_t>(tot, 1) * sizeof(int32_t)));
  for (double** p : {&c->mass, &c->stiff, &c->A, &c->B}) MONO_CUDA(c, cudaMalloc(p, std::max<int64_t>(tot, 1) * sizeof(double)));
  MONO_CUDA(c, cudaMemcpyAsync(c->slice_ptr, sp.data(), (ns + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->cols, hc.data(), tot * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->mass, hm.data(), tot * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->stiff, hk.data(), tot * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  // stencil dictionary: on by default (measured r02d: PDE stage -6 % at 3.4 M rows, -18 % at 27 M, -22 % at 0.44 M streamed;
  // bit-identical results); MONO_PDE_DICT=0 keeps every row on the SELL stream.  It only ever serves the streaming KSPCG
  // kernel, and only when it covers at least half of the rows (structured meshes with constant coefficients).
  const char* dict_env = getenv("MONO_PDE_DICT");
  if (n > 0 && !(dict_env != nullptr && dict_env[0] == '0')) return pde_build_dictionary(c, indptr, indices, mass, stiff, sp);
  return MONO_OK;
}

int pde_update_matrices(mono_ctx* c, double dt) {
  const int64_t nnz = c->sell_nnz;
  if (nnz > 0) {
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((nnz + threads - 1) / threads, (int64_t)c->n_sm * 16);
    build_ab_kernel<<<blocks, threads, 0, c->stream>>>(nnz, c->mass, c->stiff, c->A, c->B, c->C_m, dt * c->theta,
                                                       dt * (1.0 - c->theta));
    c->launches++;
    const int jb = (int)std::min<int64_t>((c->n_owned + threads - 1) / threads, (int64_t)c->n_sm * 16);
    MONO_CUDA(c, cudaMemsetAsync(c->gen_state + 2, 0, sizeof(unsigned long long), c->stream));
    jacobi_kernel<<<std::max(jb, 1), threads, 0, c->stream>>>(c->n_owned, c->n_slices, c->slice_ptr, c->cols, c->A,
                                                             c->dinv, c->pc_type == MONO_PC_CHEBYSHEV ? MONO_PC_JACOBI : c->pc_type,
                                                             c->gen_state + 2);
    c->launches++;
    if (c->n_pat > 0) {
      dict_gather_kernel<<<1, 1024, 0, c->stream>>>(c->n_pat * kChunk, c->dict_src_dev, c->A, c->B, c->dict_A_dev, c->dict_B_dev);
      c->launches++;
      dict_dinv_kernel<<<1, 64, 0, c->stream>>>(c->n_pat, c->dict_rep_dev, c->dinv, c->dict_dinv_dev);  // (after jacobi_kernel)
      c->launches++;
      if (c->nd_src_dev != nullptr && c->n_nd > 0) {
        const int64_t m = c->n_nd * kChunk;
        dict_gather_kernel<<<(int)std::min<int64_t>((m + 1023) / 1024, 4096), 1024, 0, c->stream>>>((int)m, c->nd_src_dev, c->A, c->B,
                                                                                                c->nd_A_dev, c->nd_B_dev);
        c->launches++;
      }
    }
    MONO_CUDA(c, cudaGetLastError());
  }
  if (c->pc_type == MONO_PC_CHEBYSHEV) {
    double g = 0.0;
    MONO_CUDA(c, cudaMemcpyAsync(&g, c->gen_state + 2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    int rc = halo_allreduce_max(c, &g);  // one polynomial for the whole (distributed) operator
    if (rc) return rc;
    if (!(g > 0.0)) g = 1.0;
    // Chebyshev iteration on [b/kappa, b] (Saad, Iterative Methods, Alg. 12.1), b with a hair of head room
    const double b = g * (1.0 + 1e-12), lo = b / c->cheb_kappa;
    const double theta = 0.5 * (b + lo), delta = 0.5 * (b - lo), sigma = theta / delta;
    c->cheb_inv_theta = 1.0 / theta;
    double rho = 1.0 / sigma;
    for (int j = 1; j < kMaxCheb; ++j) {
      const double rho_n = 1.0 / (2.0 * sigma - rho);
      c->cheb_c1[j] = rho_n * rho;
      c->cheb_c2[j] = 2.0 * rho_n / delta;
      rho = rho_n;
    }
    c->cheb_c1[0] = c->cheb_c2[0] = 0.0;
    c->gershgorin = g;
  }
  c->cur_dt = dt;
  c->have_dt = true;
  return MONO_OK;
}

template <class K>
static bool opt_in_smem(K kernel, size_t bytes) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess) return true;
  (void)cudaGetLastError();
  return false;
}

int pde_setup_launch_config(mono_ctx* c) {
  // worker CTAs own rows (one CTA per SM: 16 warps, 16 entries per row in flight each); one more CTA only
  // reduces the dot products
  // (slices are dealt round-robin over the worker CTAs, so a small mesh spreads over all SMs instead of filling
  // the first few: 58 176 rows = 1818 slices = 12-13 warps on each of 147 SMs)
  c->pde_workers = (int)std::max<int64_t>(1, std::min<int64_t>(c->n_slices, std::min(c->n_sm, kMaxBlocks) - 1));
  if (const char* e = getenv("MONO_PDE_WORKERS")) c->pde_workers = std::max(1, std::min(atoi(e), std::min(c->n_sm, kMaxBlocks) - 1));
  c->pde_blocks = c->pde_workers + 1;
  c->pde_threads = kPdeThreads;
  c->rows_per_thread = (int)std::max<int64_t>(
      1, (c->n_slices + (int64_t)c->pde_workers * kWarpsPerBlock - 1) / ((int64_t)c->pde_workers * kWarpsPerBlock));
  if (c->recs) cudaFree(c->recs);
  const size_t nrec = (size_t)8 * c->pde_workers + 8;  // partial sums [2][4][workers] + totals [2][4]
  MONO_CUDA(c, cudaMalloc(&c->recs, sizeof(SyncRec) * nrec));
  MONO_CUDA(c, cudaMemsetAsync(c->recs, 0, sizeof(SyncRec) * nrec, c->stream));
  // ONE allocation holds everything another rank may write into (it is exported with CUDA IPC, halo.cu):
  //   [ t0 | t1 ]  the exchanged vector, two tagged buffers over owned + ghost dofs (tag 0 = never written)
  //   [ xg ]       landing zone for the neighbours' final x values, one tagged record per ghost
  //   [ xrecs ]    cross-rank reduction records [2 parities][4 slots][kMaxRanks]
  const int64_t nl = std::max<int64_t>(c->n_local, 32);
  const int64_t ng = std::max<int64_t>(c->n_ghost, 1);
  const int nb = c->pc_type == MONO_PC_CHEBYSHEV ? 3 * c->cheb_k : 2;  // exchange buffers (pde_pipecg_kernel)
  c->exch_nb = nb;
  c->exch_nl = nl;
  c->exch_off_xg = nb * nl;
  c->exch_off_xrecs = nb * nl + ng;
  c->exch_recs = nb * nl + ng + 8 * kMaxRanks;
  if (c->exch) {
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->exch);
  }
  MONO_CUDA(c, cudaMalloc(&c->exch, sizeof(SyncRec) * c->exch_recs));
  MONO_CUDA(c, cudaMemsetAsync(c->exch, 0, sizeof(SyncRec) * c->exch_recs, c->stream));
  c->xg = c->exch + c->exch_off_xg;
  c->xrecs = c->exch + c->exch_off_xrecs;
  if (!c->gen_state) {
    MONO_CUDA(c, cudaMalloc(&c->gen_state, sizeof(unsigned long long) * 4));
    const unsigned long long init[4] = {1ull, 1ull, 0ull, 0ull};  // generation / tag 0 = "never written"
    MONO_CUDA(c, cudaMemcpyAsync(c->gen_state, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  if (const char* e = getenv("MONO_SPIN_TIMEOUT_MS")) c->spin_timeout_ms = std::max(1.0, atof(e));
  c->mode_dirty = true;
  return MONO_OK;
}

template <bool MULTI>
static const void* pde_kernel_tbl(bool pipe, bool cheb, bool resident, bool matsmem) {
  if (!pipe) {
    if (matsmem) return (const void*)pde_cg_kernel<true, true, MULTI>;
    if (resident) return (const void*)pde_cg_kernel<true, false, MULTI>;
    return (const void*)pde_cg_kernel<false, false, MULTI>;
  }
  if (matsmem) return cheb ? (const void*)pde_pipecg_kernel<true, true, MULTI, true> : (const void*)pde_pipecg_kernel<true, true, MULTI, false>;
  if (resident) return cheb ? (const void*)pde_pipecg_kernel<true, false, MULTI, true> : (const void*)pde_pipecg_kernel<true, false, MULTI, false>;
  return cheb ? (const void*)pde_pipecg_kernel<false, false, MULTI, true> : (const void*)pde_pipecg_kernel<false, false, MULTI, false>;
}

static size_t pde_ring_smem(const mono_ctx* c) {
  return ((size_t)(kRingTableSmem + 127) / 128) * 128 + ((size_t)c->ring_ncl << c->ring_cap_log2) * 8 + 2 * kRingStages * 8 +
         (size_t)kRingStages * kRingStageBytes + 64;
}

static const void* pde_kernel_for_mode(const mono_ctx* c, bool resident, bool matsmem, bool multi) {
  const bool pipe = c->ksp_type == MONO_KSP_PIPECG;
  const bool cheb = pipe && c->pc_type == MONO_PC_CHEBYSHEV && c->cheb_k > 1;
  return multi ? pde_kernel_tbl<true>(pipe, cheb, resident, matsmem) : pde_kernel_tbl<false>(pipe, cheb, resident, matsmem);
}

// Where the CG vectors (and the A rows) of this mesh live, for the configured Krylov driver and preconditioner.
static int pde_select_mode(mono_ctx* c) {
  const bool pipe = c->ksp_type == MONO_KSP_PIPECG;
  const bool cheb = pipe && c->pc_type == MONO_PC_CHEBYSHEV && c->cheb_k > 1;
  const int kind = !pipe ? 0 : cheb ? 2 : 1;
  const size_t per_row = (size_t)kNvec[kind] * kPdeThreads * sizeof(double);
  const size_t mat_bytes = (size_t)kChunk * kPdeThreads * (sizeof(double) + sizeof(int32_t));
  const size_t budget = 226 * 1024;  // 227 KB per CTA minus the kernels' static shared memory
  const bool multi = c->nranks > 1;
  const void* k = pde_kernel_for_mode(c, true, false, multi);
  c->resident = (size_t)c->rows_per_thread * per_row <= budget && getenv("MONO_PDE_STREAM") == nullptr;
  c->matsmem = c->resident && c->rows_per_thread == 1 && c->max_width <= kChunk && per_row + mat_bytes <= budget &&
               getenv("MONO_PDE_NO_MATSMEM") == nullptr;
  c->resident_smem = c->resident ? c->rows_per_thread * per_row + (c->matsmem ? mat_bytes : 0) : 0;
  c->staged = false;
  if (c->resident) {
    k = pde_kernel_for_mode(c, true, c->matsmem, multi);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->resident_smem) != cudaSuccess) {
      (void)cudaGetLastError();
      c->resident = c->matsmem = false;
      c->resident_smem = 0;
    }
  }
  // streaming KSPCG: plain vectors + fenced barrier (pde_cg_stream_kernel); MONO_PDE_TAGGED_STREAM=1 keeps the tagged
  // exchange of the resident kernels for it (measurement / fallback)
  c->stream_plain = !c->resident && !pipe && getenv("MONO_PDE_TAGGED_STREAM") == nullptr;
  c->dict_active = !c->resident && !pipe && c->n_pat > 0;
  c->ring_active = false;
  if (c->dict_active) {
    size_t smem = kDictSmem;
    if (c->stream_plain) {
      c->ring_active = c->ring_ok && getenv("MONO_PDE_NO_RING") == nullptr;
      if (c->ring_active) {
        k = multi ? (const void*)pde_cg_stream_kernel<true, kModeRing> : (const void*)pde_cg_stream_kernel<false, kModeRing>;
        smem = pde_ring_smem(c);
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
          (void)cudaGetLastError();
          c->ring_active = false;
          smem = kDictSmem;
        }
      }
      if (!c->ring_active)
        k = multi ? (const void*)pde_cg_stream_kernel<true, kModeDictL1> : (const void*)pde_cg_stream_kernel<false, kModeDictL1>;
    } else {
      k = multi ? (const void*)pde_cg_kernel<false, false, true, true> : (const void*)pde_cg_kernel<false, false, false, true>;
    }
    if (!c->ring_active && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      c->dict_active = false;
    }
    if (c->ring_active) {  // side rows of each CTA's run of slices (same split as the kernel's)
      int64_t per_cta = (c->n_slices + c->pde_workers - 1) / c->pde_workers;
      per_cta = (per_cta + kRingTile / kSlice - 1) / (kRingTile / kSlice) * (kRingTile / kSlice);
      std::vector<int64_t> ptr((size_t)c->pde_workers + 1, 0);
      for (int b = 0; b <= c->pde_workers; ++b) {
        const int64_t row_lo = std::min<int64_t>((int64_t)b * per_cta, c->n_slices) * kSlice;
        ptr[(size_t)b] = std::lower_bound(c->nd_rows_host.begin(), c->nd_rows_host.end(), row_lo,
                                          [](int32_t r, int64_t v) { return (int64_t)r < v; }) - c->nd_rows_host.begin();
      }
      if (c->nd_cta_ptr_dev) cudaFree(c->nd_cta_ptr_dev);
      c->nd_cta_ptr_dev = nullptr;
      MONO_CUDA(c, cudaMalloc(&c->nd_cta_ptr_dev, ptr.size() * sizeof(int64_t)));
      MONO_CUDA(c, cudaMemcpyAsync(c->nd_cta_ptr_dev, ptr.data(), ptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
      MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    }
  }
  if (!c->dict_active && !c->resident && !pipe && c->max_width <= kChunk && getenv("MONO_PDE_NO_STAGING") == nullptr) {
    if (c->stream_plain)
      k = multi ? (const void*)pde_cg_stream_kernel<true, kModeSell> : (const void*)pde_cg_stream_kernel<false, kModeSell>;
    else
      k = pde_kernel_for_mode(c, false, false, multi);
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kStagedSmem) == cudaSuccess)
      c->staged = true;
    else
      (void)cudaGetLastError();
  }
  c->mode_dirty = false;
  return MONO_OK;
}

static void fill_sync_args(const mono_ctx* c, PdeArgs& a) {
  a.recs = c->recs;
  a.gen_state = c->gen_state;
  a.spin_ns = (u64)(c->spin_timeout_ms * 1e6);
  a.n_workers = c->pde_workers;
  a.nranks = c->nranks;
  a.rank = c->rank;
  a.xg = c->xg;
  for (int q = 0; q < kMaxRanks; ++q) a.peer_xrecs[q] = c->peer_xrecs[q];
  a.peer_xrecs[c->rank] = c->xrecs;
  a.send_of_row = c->send_of_row_dev;
  a.send_ents = static_cast<const SendEnt*>(c->send_ents_dev);
  a.slice_send = c->slice_send_dev;
}

static int stim_refresh(mono_ctx* c, double t_eval, int* has_stim) {
  // The dense source vector sum_k a_k(t) s_k only changes when a window opens/closes or an amplitude is
  // re-assigned; rebuild it then (before the step kernel, stream-ordered), not inside the solve.
  std::vector<std::pair<int, double>> sig;
  for (int k = 0; k < (int)c->stims_host.size(); ++k) {
    const StimDev& st = c->stims_host[k];
    if (st.amp != 0.0 && t_eval >= st.t_start && t_eval <= st.t_end && st.nnz > 0) sig.emplace_back(k, st.amp);
  }
  if (sig != c->stim_sig) {
    if (!c->stim_vec) {
      const int64_t nl = std::max<int64_t>(c->n_local, 32);
      MONO_CUDA(c, cudaMalloc(&c->stim_vec, sizeof(double) * nl));
      MONO_CUDA(c, cudaMemsetAsync(c->stim_vec, 0, sizeof(double) * nl, c->stream));
    }
    const int threads = 256;
    MONO_CUDA(c, cudaMemsetAsync(c->slice_stim_dev, 0, (size_t)c->n_slices + 64, c->stream));
    auto launch = [&](const StimDev& st, double scale) {
      const int blocks = (int)std::min<int64_t>((st.nnz + threads - 1) / threads, (int64_t)c->n_sm * 8);
      stim_scatter_kernel<<<blocks, threads, 0, c->stream>>>(st.nnz, st.idx, st.val, scale, c->stim_vec, c->slice_stim_dev);
      c->launches++;
    };
    for (auto& kv : c->stim_sig) launch(c->stims_host[kv.first], 0.0);
    for (auto& kv : sig) launch(c->stims_host[kv.first], kv.second);
    MONO_CUDA(c, cudaGetLastError());
    c->stim_sig = sig;
  }
  *has_stim = sig.empty() ? 0 : 1;
  return MONO_OK;
}

int pde_launch_step(mono_ctx* c, double t_eval, double dt) {
  if (c->mode_dirty) {
    int rcm = pde_select_mode(c);
    if (rcm) return rcm;
  }
  int has_stim = 0;
  int rc = stim_refresh(c, t_eval, &has_stim);
  if (rc) return rc;
  if (!c->resident) {  // streaming mode: thread-private vectors live in global memory (allocated on first use)
    const int64_t no = std::max<int64_t>(c->n_owned, 32);
    for (int k = 0; k < 10; ++k) {
      const bool needed = c->stream_plain ? (k == VR || k == VQ || k == VP) : true;  // KSPCG keeps r, q, p (x, d alias a.x, a.dinv)
      if (needed && !c->work[k]) {
        MONO_CUDA(c, cudaMalloc(&c->work[k], sizeof(double) * (no + 2)));  // (+2: see mono_pde_set_matrices)
        MONO_CUDA(c, cudaMemsetAsync(c->work[k], 0, sizeof(double) * (no + 2), c->stream));
      }
    }
  }
  PdeArgs a;
  a.n_owned = c->n_owned;
  a.n_local = c->n_local;
  a.n_slices = c->n_slices;
  a.slice_ptr = c->slice_ptr;
  a.cols = c->cols;
  a.A = c->A;
  a.B = c->B;
  a.dinv = c->dinv;
  a.v_prev = c->v_prev;
  a.x = c->x;
  for (int k = 0; k < 10; ++k) a.work[k] = c->work[k];
  for (int b = 0; b < kMaxTb; ++b) a.tb[b] = b < c->exch_nb ? c->exch + (int64_t)b * c->exch_nl : nullptr;
  const bool use_cheb = c->pc_type == MONO_PC_CHEBYSHEV && c->ksp_type == MONO_KSP_PIPECG;
  a.cheb_k = use_cheb ? c->cheb_k : 1;  // (k == 1 runs the plain Jacobi build: same operator up to the scale 1/theta)
  a.cheb_inv_theta = c->cheb_inv_theta;
  for (int j = 0; j < kMaxCheb; ++j) {
    a.cheb_c1[j] = c->cheb_c1[j];
    a.cheb_c2[j] = c->cheb_c2[j];
  }
  if (c->pc_type == MONO_PC_CHEBYSHEV && !use_cheb)
    return mono_fail(c, MONO_E_UNSUPPORTED, "the Chebyshev preconditioner is implemented in the pipelined driver: use ksp_type pipecg");
  a.stim_vec = c->stim_vec;
  a.has_stim = has_stim;
  a.rows_per_thread = c->rows_per_thread;
  a.staged = c->staged ? 1 : 0;
  a.dt = dt;
  a.rtol = c->rtol;
  a.atol = c->atol;
  a.max_it = c->max_it;
  a.norm_type = c->norm_type;
  a.x0_mode = c->x0_mode;
  fill_sync_args(c, a);
  a.res = c->ksp_dev;
  a.timeline = c->timeline_dev;
  a.pat = c->pat_dev;
  a.dict_A = c->dict_A_dev;
  a.dict_B = c->dict_B_dev;
  a.dict_off = c->dict_off_dev;
  a.dict_w = c->dict_w_dev;
  a.n_pat = c->n_pat;
  a.dict_cl = c->dict_cl_dev;
  a.ring_ncl = c->ring_ncl;
  a.ring_cap_log2 = c->ring_cap_log2;
  a.ring_cl0 = 0;
  a.ring_depth = std::max(c->ring_depth, kRingMinDepth);
  if (const char* e = getenv("MONO_RING_DEPTH")) a.ring_depth = std::max(1, std::min(atoi(e), c->ring_depth));
  for (int q = 0; q < kMaxClusters; ++q) {
    a.ring_lo[q] = c->ring_lo[q];
    a.ring_hi[q] = c->ring_hi[q];
    if (q < c->ring_ncl && c->ring_lo[q] <= 0 && 0 <= c->ring_hi[q]) a.ring_cl0 = q;
  }
  a.nd_rows = c->nd_rows_dev;
  a.nd_w = c->nd_w_dev;
  a.nd_cols = c->nd_cols_dev;
  a.nd_A = c->nd_A_dev;
  a.nd_B = c->nd_B_dev;
  a.nd_cta_ptr = c->nd_cta_ptr_dev;
  a.dict_dinv = c->dict_dinv_dev;
  a.slice_stim = c->slice_stim_dev;
  a.dbg = 0;
  if (const char* e = getenv("MONO_RING_DBG")) a.dbg = atoi(e);
  const bool multi = c->nranks > 1;
  if (multi && !c->peers_ready)
    return mono_fail(c, MONO_E_INVALID, "multi-rank context: call mono_set_halo (on every rank) before stepping the PDE stage");
  void* args[] = {&a};
  const void* kern = pde_kernel_for_mode(c, c->resident, c->matsmem, multi);
  size_t smem = a.staged ? (size_t)kStagedSmem : c->resident_smem;
  if (c->dict_active) {
    kern = multi ? (const void*)pde_cg_kernel<false, false, true, true> : (const void*)pde_cg_kernel<false, false, false, true>;
    smem = kDictSmem;
  }
  if (c->stream_plain) {
    if (c->ring_active) {
      kern = multi ? (const void*)pde_cg_stream_kernel<true, kModeRing> : (const void*)pde_cg_stream_kernel<false, kModeRing>;
      smem = pde_ring_smem(c);
    } else if (c->dict_active) {
      kern = multi ? (const void*)pde_cg_stream_kernel<true, kModeDictL1> : (const void*)pde_cg_stream_kernel<false, kModeDictL1>;
    } else {
      kern = multi ? (const void*)pde_cg_stream_kernel<true, kModeSell> : (const void*)pde_cg_stream_kernel<false, kModeSell>;
    }
  }
  MONO_CUDA(c, cudaLaunchCooperativeKernel(kern, dim3(c->pde_blocks), dim3(c->pde_threads), args, smem, c->stream));
  c->launches++;
  return MONO_OK;
}

int pde_bench_sync(mono_ctx* c, int n, float* us_per_sync) {
  if (!c->recs) return mono_fail(c, MONO_E_INVALID, "set matrices first (the synchronisation records belong to the PDE stage)");
  PdeArgs a{};
  a.timeline = nullptr;
  fill_sync_args(c, a);
  double* out = nullptr;
  MONO_CUDA(c, cudaMalloc(&out, sizeof(double)));
  cudaEvent_t e0, e1;
  MONO_CUDA(c, cudaEventCreate(&e0));
  MONO_CUDA(c, cudaEventCreate(&e1));
  void* args[] = {&a, &n, &out};
  MONO_CUDA(c, cudaEventRecord(e0, c->stream));
  MONO_CUDA(c, cudaLaunchCooperativeKernel((void*)sync_bench_kernel, dim3(c->pde_blocks), dim3(c->pde_threads), args, 0, c->stream));
  MONO_CUDA(c, cudaEventRecord(e1, c->stream));
  MONO_CUDA(c, cudaEventSynchronize(e1));
  float ms = 0.f;
  MONO_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  c->launches++;
  if (us_per_sync) *us_per_sync = ms * 1e3f / (float)n;
  return MONO_OK;
}

// Send table of the in-kernel halo exchange: entry i says that owned row row[i] is a ghost on some neighbour
// rank whose (peer-mapped) slots in its two exchange buffers and its x landing zone are dst_*[i].
int pde_build_send_table(mono_ctx* c, const std::vector<int32_t>& row, const std::vector<std::vector<void*>>& dst_t,
                         const std::vector<void*>& dst_xg) {
  const size_t n = row.size();
  std::vector<int32_t> first((size_t)std::max<int64_t>(c->n_owned, 1), -1);
  std::vector<SendEnt> ents(std::max<size_t>(n, 1));
  for (size_t i = 0; i < n; ++i) {
    SendEnt& e = ents[i];
    for (int b = 0; b < kMaxTb; ++b) e.t[b] = b < (int)dst_t.size() ? static_cast<SyncRec*>(dst_t[b][i]) : nullptr;
    e.xg = static_cast<SyncRec*>(dst_xg[i]);
    e.next = first[row[i]];
    e.pad = 0;
    first[row[i]] = (int32_t)i;
  }
  std::vector<uint8_t> slice_send((size_t)std::max<int64_t>(c->n_slices, 1), 0);
  for (size_t i = 0; i < n; ++i) slice_send[(size_t)(row[i] / kSlice)] = 1;
  if (c->send_of_row_dev) cudaFree(c->send_of_row_dev);
  if (c->send_ents_dev) cudaFree(c->send_ents_dev);
  if (c->slice_send_dev) cudaFree(c->slice_send_dev);
  c->send_of_row_dev = nullptr;
  c->send_ents_dev = nullptr;
  c->slice_send_dev = nullptr;
  MONO_CUDA(c, cudaMalloc(&c->slice_send_dev, slice_send.size()));
  MONO_CUDA(c, cudaMemcpyAsync(c->slice_send_dev, slice_send.data(), slice_send.size(), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMalloc(&c->send_of_row_dev, sizeof(int32_t) * first.size()));
  MONO_CUDA(c, cudaMalloc(&c->send_ents_dev, sizeof(SendEnt) * ents.size()));
  MONO_CUDA(c, cudaMemcpyAsync(c->send_of_row_dev, first.data(), sizeof(int32_t) * first.size(), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->send_ents_dev, ents.data(), sizeof(SendEnt) * ents.size(), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  return MONO_OK;
}

int probes_launch(mono_ctx* c, double t0) {
  const int n = (int)c->probes_host.size();
  const bool field = c->actmap_enabled || c->minmax_enabled;
  if (n == 0 && !field) return MONO_OK;
  if (n > 0 && c->probes_dirty) {
    if (c->probes_dev) cudaFree(c->probes_dev);
    if (c->probe_vals_dev) cudaFree(c->probe_vals_dev);
    double* old_act = c->probe_act_dev;
    MONO_CUDA(c, cudaMalloc(&c->probes_dev, sizeof(ProbeDev) * n));
    MONO_CUDA(c, cudaMalloc(&c->probe_vals_dev, sizeof(double) * n));
    MONO_CUDA(c, cudaMalloc(&c->probe_act_dev, sizeof(double) * n));
    std::vector<double> neg(n, -1.0);
    MONO_CUDA(c, cudaMemcpyAsync(c->probes_dev, c->probes_host.data(), sizeof(ProbeDev) * n, cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaMemcpyAsync(c->probe_act_dev, neg.data(), sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    if (old_act) cudaFree(old_act);
    c->probes_dirty = false;
  }
  int blocks = (n + 255) / 256;
  if (field) {
    blocks = std::max(blocks, (int)std::min<int64_t>((c->n_owned + 255) / 256, (int64_t)c->n_sm * 8));
    if (c->minmax_enabled) {  // identities of min / max in the ordered-integer image
      const unsigned long long init[2] = {~0ull, 0ull};
      MONO_CUDA(c, cudaMemcpyAsync(c->minmax_dev, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    }
  }
  observers_kernel<<<std::max(blocks, 1), 256, 0, c->stream>>>(n, c->probes_dev, c->x, c->probe_vals_dev, c->probe_act_dev,
                                                             c->act_enabled ? 1 : 0, c->act_threshold, t0, c->n_owned,
                                                             c->actmap_enabled ? c->actmap_dev : nullptr, c->actmap_threshold,
                                                             c->minmax_enabled ? c->minmax_dev : nullptr);
  c->launches++;
  MONO_CUDA(c, cudaGetLastError());
  return MONO_OK;
}

int strided_pack_launch(mono_ctx* c, int64_t offset, int64_t stride, int64_t count, double* out_dev) {
  const int blocks = (int)std::min<int64_t>((count + 255) / 256, (int64_t)c->n_sm * 8);
  strided_pack_kernel<<<std::max(blocks, 1), 256, 0, c->stream>>>(count, offset, stride, c->x, out_dev);
  c->launches++;
  MONO_CUDA(c, cudaGetLastError());
  return MONO_OK;
}
