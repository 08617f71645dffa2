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
// own column.  Algorithmic traffic: 12 B per stored entry + vectors = 200 B per row per SpMV for the
// 15-entry rows of a Kuhn-split box mesh (HBM-bound; DESIGN.md "K4").
//
// The whole solve (RHS, stimulus, initial residual, every CG iteration, convergence test) runs inside
// one cooperative launch: grid = (CTAs that fit) with a generation-counting grid barrier, two barriers
// per iteration.  Dot products are reduced per CTA, written to a partials array and summed by every CTA
// in the same fixed order after the barrier, so all CTAs take the same branch on the convergence test
// and the result is run-to-run deterministic.  No host round trip per iteration, no launch latency per
// vector operation - which is what bounds the 58k-node Niederer slab.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "mono_ctx.h"

namespace {

constexpr int kSlice = 32;
constexpr int kPdeThreads = 512;
constexpr int kWarpsPerBlock = kPdeThreads / 32;

struct PdeArgs {
  int64_t n_owned, n_local, n_slices;
  const int64_t* slice_ptr;
  const int32_t* cols;
  const double* A;
  const double* B;
  const double* dinv;
  double *x, *v_prev, *b, *r, *z, *p0, *p1, *q;
  int n_stim;
  const StimDev* stims;
  double t_eval, dt;
  double rtol, atol;
  int max_it, norm_type, x0_mode;
  unsigned* bar;
  double* partials;  // [2][4][gridDim.x]
  KspResult* res;
};

// ---- grid-wide barrier (all CTAs are co-resident: cooperative launch) ---------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned* gen = bar + 1;
    const unsigned g = *gen;
    __threadfence();
    if (atomicAdd(bar, 1u) == nblocks - 1) {
      atomicExch(bar, 0u);
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen == g) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduce NV per-thread values over the CTA and store the CTA's partial sums (slot-major) for this parity.
template <int NV>
__device__ __forceinline__ void block_partials(const double (&v)[NV], double* partials, int parity, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) smem[k * kWarpsPerBlock + warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = lane < kWarpsPerBlock ? smem[k * kWarpsPerBlock + lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) partials[((size_t)parity * 4 + k) * gridDim.x + blockIdx.x] = s;
    }
  }
}

// After the barrier: every CTA sums all CTAs' partials in the same order -> identical totals everywhere.
template <int NV>
__device__ __forceinline__ void grid_totals(double (&out)[NV], const double* partials, int parity, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp < NV) {
    const volatile double* src = partials + ((size_t)parity * 4 + warp) * gridDim.x;
    double s = 0.0;
    for (unsigned i = lane; i < gridDim.x; i += 32) s += src[i];
    s = warp_sum(s);
    if (lane == 0) smem[64 + warp] = s;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) out[k] = smem[64 + k];
  __syncthreads();
}

__global__ void __launch_bounds__(kPdeThreads, 1) pde_step_kernel(const PdeArgs a) {
  __shared__ double smem[64 + 8];
  const unsigned nb = gridDim.x;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)nb * kWarpsPerBlock;
  const int64_t tid = (int64_t)blockIdx.x * kPdeThreads + threadIdx.x;
  const int64_t tstride = (int64_t)nb * kPdeThreads;
  const bool x0_prev = a.x0_mode == MONO_X0_PREVIOUS;
  int parity = 0;

  // ---- K2: b = B v_   (and q = A v_ when the initial guess is v_) --------------------------------
  for (int64_t s = warp_global; s < a.n_slices; s += warp_stride) {
    const int64_t row = s * kSlice + lane;
    const int64_t beg = a.slice_ptr[s];
    const int width = (int)((a.slice_ptr[s + 1] - beg) / kSlice);
    double accB = 0.0, accA = 0.0;
    for (int k = 0; k < width; ++k) {
      const int64_t e = beg + (int64_t)k * kSlice + lane;
      const double vj = a.v_prev[a.cols[e]];
      accB = fma(a.B[e], vj, accB);
      if (x0_prev) accA = fma(a.A[e], vj, accA);
    }
    if (row < a.n_owned) {
      a.b[row] = accB;
      if (x0_prev) a.q[row] = accA;
    }
  }
  // ---- stimulus: b += dt * a_k(t) * s_k for every stimulus whose window contains t ----------------
  bool any_stim = false;
  for (int k = 0; k < a.n_stim; ++k) {
    const StimDev st = a.stims[k];
    if (st.amp != 0.0 && a.t_eval >= st.t_start && a.t_eval <= st.t_end && st.nnz > 0) {
      grid_barrier(a.bar, nb);  // rows of b are complete / previous stimulus applied
      any_stim = true;
      const double f = a.dt * st.amp;
      for (int64_t e = tid; e < st.nnz; e += tstride) a.b[st.idx[e]] += f * st.val[e];
    }
  }
  if (any_stim) grid_barrier(a.bar, nb);

  // ---- initial residual, z = D^-1 r, p = z --------------------------------------------------------
  {
    double acc[3] = {0.0, 0.0, 0.0};  // r.z, norm^2 of r (chosen norm), norm^2 of b (chosen norm)
    for (int64_t i = tid; i < a.n_owned; i += tstride) {
      const double bi = a.b[i];
      const double di = a.dinv[i];
      double xi, ri;
      if (x0_prev) {
        xi = a.v_prev[i];
        ri = bi - a.q[i];
      } else {
        xi = 0.0;
        ri = bi;
      }
      const double zi = di * ri;
      a.x[i] = xi;
      a.r[i] = ri;
      a.z[i] = zi;
      a.p0[i] = zi;
      acc[0] = fma(ri, zi, acc[0]);
      const double zb = di * bi;
      if (a.norm_type == MONO_NORM_PRECONDITIONED) {
        acc[1] = fma(zi, zi, acc[1]);
        acc[2] = fma(zb, zb, acc[2]);
      } else if (a.norm_type == MONO_NORM_UNPRECONDITIONED) {
        acc[1] = fma(ri, ri, acc[1]);
        acc[2] = fma(bi, bi, acc[2]);
      } else {
        acc[1] = fma(ri, zi, acc[1]);
        acc[2] = fma(bi, zb, acc[2]);
      }
    }
    block_partials<3>(acc, a.partials, parity, smem);
  }
  grid_barrier(a.bar, nb);
  double tot[3];
  grid_totals<3>(tot, a.partials, parity, smem);
  parity ^= 1;
  double rz = tot[0];
  double rnorm = sqrt(fabs(tot[1]));
  const double bnorm = sqrt(fabs(tot[2]));
  const double ttol = fmax(a.rtol * bnorm, a.atol);
  int its = 0;
  int reason = 0;
  if (!(rnorm == rnorm)) {
    reason = MONO_KSP_DIVERGED_NAN;
  } else if (rnorm <= ttol) {
    reason = rnorm <= a.atol ? MONO_KSP_CONVERGED_ATOL : MONO_KSP_CONVERGED_RTOL;
  } else if (a.max_it <= 0) {
    reason = MONO_KSP_DIVERGED_ITS;
  }

  // p_cur holds the search direction of the current iteration; for it >= 1 it is formed on the fly
  // inside the SpMV gather as z + beta*p_old (both complete vectors), and written for the own row.
  double* p_old = a.p0;
  double* p_new = a.p1;
  double beta = 0.0;
  bool first = true;
  while (reason == 0) {
    // ---- K4a: q = A p, p.q -----------------------------------------------------------------------
    double pq[1] = {0.0};
    for (int64_t s = warp_global; s < a.n_slices; s += warp_stride) {
      const int64_t row = s * kSlice + lane;
      const int64_t beg = a.slice_ptr[s];
      const int width = (int)((a.slice_ptr[s + 1] - beg) / kSlice);
      double acc = 0.0;
      if (first) {
        for (int k = 0; k < width; ++k) {
          const int64_t e = beg + (int64_t)k * kSlice + lane;
          acc = fma(a.A[e], p_old[a.cols[e]], acc);
        }
        if (row < a.n_owned) {
          a.q[row] = acc;
          pq[0] = fma(p_old[row], acc, pq[0]);
        }
      } else {
        for (int k = 0; k < width; ++k) {
          const int64_t e = beg + (int64_t)k * kSlice + lane;
          const int32_t j = a.cols[e];
          acc = fma(a.A[e], fma(beta, p_old[j], a.z[j]), acc);
        }
        if (row < a.n_owned) {
          const double pi = fma(beta, p_old[row], a.z[row]);
          p_new[row] = pi;
          a.q[row] = acc;
          pq[0] = fma(pi, acc, pq[0]);
        }
      }
    }
    block_partials<1>(pq, a.partials, parity, smem);
    grid_barrier(a.bar, nb);
    double t1[1];
    grid_totals<1>(t1, a.partials, parity, smem);
    parity ^= 1;
    if (!first) {
      double* t = p_old;
      p_old = p_new;
      p_new = t;
    }
    first = false;
    const double alpha = rz / t1[0];
    // ---- K4b: x += alpha p ; r -= alpha q ; z = D^-1 r ; r.z and the residual norm ---------------
    double acc[2] = {0.0, 0.0};
    for (int64_t i = tid; i < a.n_owned; i += tstride) {
      const double pi = p_old[i];
      const double ri = fma(-alpha, a.q[i], a.r[i]);
      const double zi = a.dinv[i] * ri;
      a.x[i] = fma(alpha, pi, a.x[i]);
      a.r[i] = ri;
      a.z[i] = zi;
      acc[0] = fma(ri, zi, acc[0]);
      if (a.norm_type == MONO_NORM_PRECONDITIONED)
        acc[1] = fma(zi, zi, acc[1]);
      else if (a.norm_type == MONO_NORM_UNPRECONDITIONED)
        acc[1] = fma(ri, ri, acc[1]);
      else
        acc[1] = fma(ri, zi, acc[1]);
    }
    block_partials<2>(acc, a.partials, parity, smem);
    grid_barrier(a.bar, nb);
    double t2[2];
    grid_totals<2>(t2, a.partials, parity, smem);
    parity ^= 1;
    ++its;
    rnorm = sqrt(fabs(t2[1]));
    beta = t2[0] / rz;
    rz = t2[0];
    if (!(rnorm == rnorm) || !(alpha == alpha)) {
      reason = MONO_KSP_DIVERGED_NAN;
    } else if (rnorm <= ttol) {
      reason = rnorm <= a.atol ? MONO_KSP_CONVERGED_ATOL : MONO_KSP_CONVERGED_RTOL;
    } else if (its >= a.max_it) {
      reason = MONO_KSP_DIVERGED_ITS;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.res->iterations = its;
    a.res->reason = reason;
    a.res->rnorm = rnorm;
    a.res->total_iterations += its;
    a.res->solves += 1;
  }
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

__global__ void jacobi_kernel(int64_t n_owned, int64_t n_slices, const int64_t* __restrict__ slice_ptr,
                              const int32_t* __restrict__ cols, const double* __restrict__ A,
                              double* __restrict__ dinv, int pc_type) {
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
    double d = 0.0;
    for (int k = 0; k < width; ++k) {
      const int64_t e = beg + (int64_t)k * kSlice + lane;
      if (cols[e] == row) d += A[e];
    }
    dinv[row] = 1.0 / d;
  }
}

__global__ void probes_kernel(int n_probes, const ProbeDev* __restrict__ probes, const double* __restrict__ x,
                              double* __restrict__ vals, double* __restrict__ act, int act_enabled,
                              double threshold, double t0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_probes) return;
  const ProbeDev pr = probes[p];
  double v = 0.0;
  for (int k = 0; k < pr.n; ++k) v = fma(pr.w[k], x[pr.node[k]], v);
  vals[p] = v;
  if (act_enabled && act[p] < 0.0 && v > threshold) act[p] = t0;
}

}  // namespace

int pde_build_sell(mono_ctx* c, const int64_t* indptr, const int32_t* indices, const double* mass, const double* stiff) {
  const int64_t n = c->n_owned;
  const int64_t ns = (n + kSlice - 1) / kSlice;
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
  MONO_CUDA(c, cudaMalloc(&c->slice_ptr, (ns + 1) * sizeof(int64_t)));
  MONO_CUDA(c, cudaMalloc(&c->cols, std::max<int64_t>(tot, 1) * sizeof(int32_t)));
  for (double** p : {&c->mass, &c->stiff, &c->A, &c->B}) MONO_CUDA(c, cudaMalloc(p, std::max<int64_t>(tot, 1) * sizeof(double)));
  MONO_CUDA(c, cudaMemcpyAsync(c->slice_ptr, sp.data(), (ns + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->cols, hc.data(), tot * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->mass, hm.data(), tot * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaMemcpyAsync(c->stiff, hk.data(), tot * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
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
    jacobi_kernel<<<std::max(jb, 1), threads, 0, c->stream>>>(c->n_owned, c->n_slices, c->slice_ptr, c->cols, c->A,
                                                             c->dinv, c->pc_type);
    c->launches++;
    MONO_CUDA(c, cudaGetLastError());
  }
  c->cur_dt = dt;
  c->have_dt = true;
  return MONO_OK;
}

int pde_setup_launch_config(mono_ctx* c) {
  int per_sm = 0;
  MONO_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pde_step_kernel, kPdeThreads, 0));
  if (per_sm < 1) return mono_fail(c, MONO_E_CUDA, "pde_step_kernel does not fit on an SM");
  // one CTA per SM (16 warps): enough loads in flight for HBM, cheapest barrier
  const int64_t need = (c->n_slices + kWarpsPerBlock - 1) / kWarpsPerBlock;
  c->pde_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(need, c->n_sm));
  c->pde_threads = kPdeThreads;
  if (c->partials) cudaFree(c->partials);
  MONO_CUDA(c, cudaMalloc(&c->partials, sizeof(double) * 2 * 4 * c->pde_blocks));
  MONO_CUDA(c, cudaMemsetAsync(c->partials, 0, sizeof(double) * 2 * 4 * c->pde_blocks, c->stream));
  return MONO_OK;
}

int pde_launch_step(mono_ctx* c, double t_eval, double dt) {
  if (c->stims_dirty) {
    const int n = (int)c->stims_host.size();
    if (n > c->stims_dev_cap) {
      if (c->stims_dev) cudaFree(c->stims_dev);
      c->stims_dev_cap = std::max(n, 8);
      MONO_CUDA(c, cudaMalloc(&c->stims_dev, sizeof(StimDev) * c->stims_dev_cap));
    }
    if (n > 0)
      MONO_CUDA(c, cudaMemcpyAsync(c->stims_dev, c->stims_host.data(), sizeof(StimDev) * n, cudaMemcpyHostToDevice, c->stream));
    c->stims_dirty = false;
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
  a.x = c->x;
  a.v_prev = c->v_prev;
  a.b = c->b;
  a.r = c->r;
  a.z = c->z;
  a.p0 = c->p0;
  a.p1 = c->p1;
  a.q = c->q;
  a.n_stim = (int)c->stims_host.size();
  a.stims = c->stims_dev;
  a.t_eval = t_eval;
  a.dt = dt;
  a.rtol = c->rtol;
  a.atol = c->atol;
  a.max_it = c->max_it;
  a.norm_type = c->norm_type;
  a.x0_mode = c->x0_mode;
  a.bar = c->bar;
  a.partials = c->partials;
  a.res = c->ksp_dev;
  void* args[] = {&a};
  MONO_CUDA(c, cudaLaunchCooperativeKernel((void*)pde_step_kernel, dim3(c->pde_blocks), dim3(c->pde_threads), args, 0, c->stream));
  c->launches++;
  return MONO_OK;
}

int probes_launch(mono_ctx* c, double t0) {
  const int n = (int)c->probes_host.size();
  if (n == 0) return MONO_OK;
  if (c->probes_dirty) {
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
  probes_kernel<<<(n + 63) / 64, 64, 0, c->stream>>>(n, c->probes_dev, c->x, c->probe_vals_dev, c->probe_act_dev,
                                                   c->act_enabled ? 1 : 0, c->act_threshold, t0);
  c->launches++;
  MONO_CUDA(c, cudaGetLastError());
  return MONO_OK;
}
