// K1: the pointwise ionic-model update, one fp64 kernel per step.
//
// Replaces the NumPy ufunc chain the reference runs in ODESystemSolver.step
// (src/beat/odesolver.py:67-79: states[:] = fun(states, t, parameters, dt)) plus the V hand-off copies
// around it (odesolver.py:164-170, utils.py:52-54, monodomain_model.py:59-60).
//
// Layout: states is (num_states, ld) row-major = structure of arrays; thread i owns node i, so every
// state row is read and written with perfectly coalesced 8-byte accesses (a warp moves 256 contiguous
// bytes per row).  All states live in registers for the whole update; shared parameters arrive in the
// kernel-parameter constant bank (so the compiler folds them into DFMA operands) together with the
// parameter-only derived constants that the host evaluated once.
//
// Roofline: fp64 pipe / MUFU bound, not HBM: 304 B of state traffic per TP06 node-step against
// ~2.9k fp64 instructions (DESIGN.md section "K1").
#include <cstdio>
#include <cstdlib>

#include "mono_ctx.h"

#ifndef MONO_ODE_MATH
#define MONO_ODE_MATH 1
#endif
#include "generated/fhn.cuh"
#include "generated/simple.cuh"
#include "generated/torord.cuh"
#include "generated/tp06.cuh"

namespace {

template <int NP, int ND>
struct UniformParams {
  static constexpr bool kPerNode = false;
  double v[NP + (ND > 0 ? ND : 1)];
  template <int K>
  __device__ __forceinline__ double p() const {
    return v[K];
  }
  template <int K>
  __device__ __forceinline__ double u() const {
    return v[NP + K];
  }
};

struct NodeParams {
  static constexpr bool kPerNode = true;
  const double* base;
  int64_t ld;
  int64_t i;
  template <int K>
  __device__ __forceinline__ double p() const {
    return __ldg(base + (int64_t)K * ld + i);
  }
  template <int K>
  __device__ __forceinline__ double u() const {
    return 0.0;
  }
};

// Parameter set chosen per node by a region index (DolfinMultiODESolver, src/beat/odesolver.py:228-354): a small table
// [n_regions][NP + ND] of shared parameters and their host-evaluated derived constants in global memory; a warp whose
// nodes share a region reads it as a broadcast out of L1.
template <int NP>
struct RegionParams {
  static constexpr bool kPerNode = false;
  const double* row;
  template <int K>
  __device__ __forceinline__ double p() const {
    return __ldg(row + K);
  }
  template <int K>
  __device__ __forceinline__ double u() const {
    return __ldg(row + NP + K);
  }
};

struct OdeArgs {
  double* states;
  int64_t ld;
  int64_t n;
  int vidx;
  const double* v_in;
  double* v_out1;
  double* v_out2;
  double t, dt;
};

#define MONO_STEP_FN(NAME, FN, NS_)                                                         \
  struct NAME {                                                                             \
    static constexpr int NS = NS_;                                                          \
    template <class PRM>                                                                    \
    __device__ __forceinline__ static void run(double (&y)[NS_], const PRM& prm, double t, double dt) { \
      FN(y, prm, t, dt);                                                                    \
    }                                                                                       \
  };
MONO_STEP_FN(fhn_fe_fn, fhn_fe, 2)
MONO_STEP_FN(fhn_grl1_fn, fhn_grl1, 2)
MONO_STEP_FN(simple_fe_fn, simple_fe, 2)
MONO_STEP_FN(simple_grl1_fn, simple_grl1, 2)
MONO_STEP_FN(tp06_fe_fn, tp06_fe, 19)
MONO_STEP_FN(tp06_grl1_fn, tp06_grl1, 19)
MONO_STEP_FN(torord_fe_fn, torord_fe, 45)
MONO_STEP_FN(torord_grl1_fn, torord_grl1, 45)

constexpr int kOdeThreads = 128;

template <class STEP, class PRM>
__device__ __forceinline__ void ode_node(const OdeArgs& a, const PRM& prm, int64_t i) {
  constexpr int NS = STEP::NS;
  double y[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) y[k] = a.states[(int64_t)k * a.ld + i];
  if (a.v_in != nullptr) {
    const double v = a.v_in[i];
#pragma unroll
    for (int k = 0; k < NS; ++k)
      if (k == a.vidx) y[k] = v;
  }
  STEP::run(y, prm, a.t, a.dt);
  double vnew = 0.0;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    a.states[(int64_t)k * a.ld + i] = y[k];
    if (k == a.vidx) vnew = y[k];
  }
  if (a.v_out1 != nullptr) a.v_out1[i] = vnew;
  if (a.v_out2 != nullptr) a.v_out2[i] = vnew;
}

template <class STEP, class UPRM>
__global__ void __launch_bounds__(kOdeThreads) ode_kernel_uniform(const OdeArgs a, const __grid_constant__ UPRM prm) {
  const int64_t i = (int64_t)blockIdx.x * kOdeThreads + threadIdx.x;
  if (i < a.n) ode_node<STEP>(a, prm, i);
}

// Same kernel compiled for 13 one-warp CTAs per SM (<= 152 registers per thread): 416 nodes per SM in flight instead
// of 384.  The kernel is latency bound at small N (every thread is one long dependent fp64 chain), so what matters
// is the number of WAVES: the BASELINE slab has 58 176 nodes = 393 per SM, one wave here and two with the
// 166-register build above.  ode_launch picks the variant with fewer waves.
constexpr int kOdeThreadsSmall = 32;
constexpr int kOdeSmallBlocksPerSm = 13;
template <class STEP, class UPRM>
__global__ void __launch_bounds__(kOdeThreadsSmall, kOdeSmallBlocksPerSm)
    ode_kernel_uniform_small(const OdeArgs a, const __grid_constant__ UPRM prm) {
  const int64_t i = (int64_t)blockIdx.x * kOdeThreadsSmall + threadIdx.x;
  if (i < a.n) ode_node<STEP>(a, prm, i);
}

template <class STEP>
__global__ void __launch_bounds__(kOdeThreads) ode_kernel_pernode(const OdeArgs a, const double* params, int64_t ldp) {
  const int64_t i = (int64_t)blockIdx.x * kOdeThreads + threadIdx.x;
  if (i < a.n) {
    NodeParams prm{params, ldp, i};
    ode_node<STEP>(a, prm, i);
  }
}

template <class K>
int blocks_per_sm(K kernel, int threads) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess) {
    (void)cudaGetLastError();
    nb = 1;
  }
  return nb > 0 ? nb : 1;
}

template <class STEP, int NP, int ND>
__global__ void __launch_bounds__(kOdeThreads) ode_kernel_regions(const OdeArgs a, const double* __restrict__ table,
                                                                    const int32_t* __restrict__ region) {
  const int64_t i = (int64_t)blockIdx.x * kOdeThreads + threadIdx.x;
  if (i < a.n) {
    RegionParams<NP> prm{table + (size_t)__ldg(region + i) * (NP + (ND > 0 ? ND : 1))};
    ode_node<STEP>(a, prm, i);
  }
}

template <class STEP, class META>
int launch_model(mono_ctx* c, const OdeArgs& a) {
  const unsigned grid = (unsigned)((a.n + kOdeThreads - 1) / kOdeThreads);
  if (grid == 0) return MONO_OK;
  if (c->region_of_node != nullptr) {
    ode_kernel_regions<STEP, META::kNumParams, META::kNumDerived><<<grid, kOdeThreads, 0, c->stream>>>(a, c->region_table, c->region_of_node);
  } else if (c->per_node) {
    ode_kernel_pernode<STEP><<<grid, kOdeThreads, 0, c->stream>>>(a, c->params_dev, c->ld);
  } else {
    using UPRM = UniformParams<META::kNumParams, META::kNumDerived>;
    UPRM prm;
    const int total = META::kNumParams + META::kNumDerived;
    for (int k = 0; k < total; ++k) prm.v[k] = k < (int)c->params_host.size() ? c->params_host[k] : 0.0;
    bool small = false;
    if constexpr (META::kNumStates <= 24) {  // the register cap only pays for models that fit it without heavy spilling
      static const int occ_big = blocks_per_sm(ode_kernel_uniform<STEP, UPRM>, kOdeThreads);
      static const int occ_small = blocks_per_sm(ode_kernel_uniform_small<STEP, UPRM>, kOdeThreadsSmall);
      const int64_t gs = (a.n + kOdeThreadsSmall - 1) / kOdeThreadsSmall;
      const int64_t waves_big = ((int64_t)grid + (int64_t)c->n_sm * occ_big - 1) / ((int64_t)c->n_sm * occ_big);
      const int64_t waves_small = (gs + (int64_t)c->n_sm * occ_small - 1) / ((int64_t)c->n_sm * occ_small);
      small = waves_small < waves_big && waves_small <= 2;
      if (const char* e = getenv("MONO_ODE_VARIANT")) small = e[0] == 's';  // measurement: force the variant (b / s)
      if (small) ode_kernel_uniform_small<STEP, UPRM><<<(unsigned)gs, kOdeThreadsSmall, 0, c->stream>>>(a, prm);
    }
    if (!small) ode_kernel_uniform<STEP, UPRM><<<grid, kOdeThreads, 0, c->stream>>>(a, prm);
  }
  c->launches++;
  MONO_CUDA(c, cudaGetLastError());
  return MONO_OK;
}

}  // namespace

int ode_model_dims(int model_id, int* ns, int* np) {
  switch (model_id) {
    case MONO_MODEL_FHN: *ns = fhn_meta::kNumStates; *np = fhn_meta::kNumParams; return 0;
    case MONO_MODEL_SIMPLE: *ns = simple_meta::kNumStates; *np = simple_meta::kNumParams; return 0;
    case MONO_MODEL_TP06: *ns = tp06_meta::kNumStates; *np = tp06_meta::kNumParams; return 0;
    case MONO_MODEL_TORORD: *ns = torord_meta::kNumStates; *np = torord_meta::kNumParams; return 0;
  }
  return -1;
}

int ode_model_num_derived(int model_id, int /*scheme_id*/) {
  switch (model_id) {
    case MONO_MODEL_FHN: return fhn_meta::kNumDerived;
    case MONO_MODEL_SIMPLE: return simple_meta::kNumDerived;
    case MONO_MODEL_TP06: return tp06_meta::kNumDerived;
    case MONO_MODEL_TORORD: return torord_meta::kNumDerived;
  }
  return -1;
}

int ode_launch(mono_ctx* c, double t, double dt, const double* v_in, double* v_out1, double* v_out2) {
  OdeArgs a{c->states, c->ld, c->npts, c->v_index, v_in, v_out1, v_out2, t, dt};
  const bool grl = c->scheme_id == MONO_SCHEME_GRL1;
  switch (c->model_id) {
    case MONO_MODEL_FHN:
      return grl ? launch_model<fhn_grl1_fn, fhn_meta>(c, a) : launch_model<fhn_fe_fn, fhn_meta>(c, a);
    case MONO_MODEL_SIMPLE:
      return grl ? launch_model<simple_grl1_fn, simple_meta>(c, a) : launch_model<simple_fe_fn, simple_meta>(c, a);
    case MONO_MODEL_TP06:
      return grl ? launch_model<tp06_grl1_fn, tp06_meta>(c, a) : launch_model<tp06_fe_fn, tp06_meta>(c, a);
    case MONO_MODEL_TORORD:
      return grl ? launch_model<torord_grl1_fn, torord_meta>(c, a) : launch_model<torord_fe_fn, torord_meta>(c, a);
  }
  return mono_fail(c, MONO_E_INVALID, "unknown model id");
}
