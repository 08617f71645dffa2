// Micro-benchmark of grid-wide synchronisation + 3-scalar reduction variants (one 512-thread CTA per SM,
// cooperative launch).  Prints microseconds per synchronisation.  Build: nvcc -arch=sm_100a -O3 -o sync_bench sync_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
struct alignas(16) Rec { double val; unsigned long long gen; };
constexpr int T = 512, W = T / 32;
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int MODE> __device__ __forceinline__ void st_rec(Rec* p, double v, unsigned long long g) {
  if (MODE == 0) asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(g) : "memory");
  if (MODE == 1) asm volatile("st.release.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(g) : "memory");
  if (MODE == 2) asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(g) : "memory");
}
template <int MODE> __device__ __forceinline__ void ld_rec(const Rec* p, double& v, unsigned long long& g) {
  long long a;
  if (MODE == 0) asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(g) : "l"(p) : "memory");
  if (MODE == 1) asm volatile("ld.acquire.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(g) : "l"(p) : "memory");
  if (MODE == 2) asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(g) : "l"(p) : "memory");
  v = __longlong_as_double(a);
}
// MODE 0: volatile + __threadfence both sides ; 1: release/acquire, no fences ; 2: relaxed + fence.acq_rel.gpu
template <int MODE, int NV> __device__ void allreduce(double (&v)[NV], Rec* recs, unsigned long long gen, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nb = gridDim.x;
  const int parity = (int)(gen & 1ull);
  for (int k = 0; k < NV; ++k) { const double s = warp_sum(v[k]); if (lane == 0) smem[k * W + warp] = s; }
  __syncthreads();
  if (warp == 0) {
    double mine[NV];
    for (int k = 0; k < NV; ++k) { double s = lane < W ? smem[k * W + lane] : 0.0; mine[k] = warp_sum(s); }
    if (lane == 0) {
      if (MODE == 0) __threadfence();
      if (MODE == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      for (int k = 0; k < NV; ++k) st_rec<MODE>(recs + ((size_t)parity * 4 + k) * nb + blockIdx.x, mine[k], gen);
    }
  }
  if (warp < NV) {
    const Rec* src = recs + ((size_t)parity * 4 + warp) * nb;
    double s = 0.0;
    for (unsigned i = lane; i < nb; i += 32) {
      double val; unsigned long long g;
      do { ld_rec<MODE>(src + i, val, g); } while (g != gen);
      s += val;
    }
    if (MODE == 0) __threadfence();
    if (MODE == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    s = warp_sum(s);
    if (lane == 0) smem[64 + warp] = s;
  }
  __syncthreads();
  for (int k = 0; k < NV; ++k) v[k] = smem[64 + k];
  __syncthreads();
}
// MODE 5/6: arrival counter; the last CTA to arrive sums all partials in fixed order and broadcasts {total, gen}.
//   5: atom.add.acq_rel.gpu + relaxed polling + one fence after ; 6: same but explicit fences around a relaxed atomic
template <int MODE, int NV> __device__ void allreduce_last(double (&v)[NV], double* partials, unsigned* counters, Rec* bc, unsigned long long gen, double* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nb = gridDim.x;
  const int parity = (int)(gen & 1ull);
  for (int k = 0; k < NV; ++k) { const double s = warp_sum(v[k]); if (lane == 0) smem[k * W + warp] = s; }
  __syncthreads();
  if (warp == 0) {
    double mine[NV];
    for (int k = 0; k < NV; ++k) { double s = lane < W ? smem[k * W + lane] : 0.0; mine[k] = warp_sum(s); }
    unsigned old = 0;
    if (lane == 0) {
      for (int k = 0; k < NV; ++k) __stcg(partials + ((size_t)parity * 4 + k) * nb + blockIdx.x, mine[k]);
      if (MODE == 5) asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(counters + parity) : "memory");
      if (MODE == 6) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); old = atomicAdd(counters + parity, 1u); asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
    }
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old == nb - 1) {  // last arriver: reduce + broadcast
      double tot[NV];
      for (int k = 0; k < NV; ++k) {
        double s = 0.0;
        for (unsigned i = lane; i < nb; i += 32) s += __ldcg(partials + ((size_t)parity * 4 + k) * nb + i);
        tot[k] = warp_sum(s);
      }
      if (lane == 0) {
        counters[parity] = 0;
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        for (int k = 0; k < NV; ++k) st_rec<2>(bc + parity * 4 + k, tot[k], gen);
      }
    }
    if (lane < NV) {
      double val; unsigned long long g;
      do { ld_rec<2>(bc + parity * 4 + lane, val, g); } while (g != gen);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      smem[64 + lane] = val;
    }
  }
  __syncthreads();
  for (int k = 0; k < NV; ++k) v[k] = smem[64 + k];
  __syncthreads();
}
// MODE 7: fence-free.  Every CTA publishes tagged records; CTA 0 polls them all (one record per thread), sums in
// fixed order and publishes tagged totals; everyone polls the totals.  No fence, no atomic.
template <int NV> __device__ void allreduce_tagged(double (&v)[NV], Rec* recs, Rec* bc, unsigned long long gen, double* smem, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nb = gridDim.x;
  const int parity = (int)(gen & 1ull);
  for (int k = 0; k < NV; ++k) { const double s = warp_sum(v[k]); if (lane == 0) smem[k * W + warp] = s; }
  __syncthreads();
  if (warp == 0) {
    double mine[NV];
    for (int k = 0; k < NV; ++k) { double s = lane < W ? smem[k * W + lane] : 0.0; mine[k] = warp_sum(s); }
    if (lane == 0) for (int k = 0; k < NV; ++k) st_rec<2>(recs + ((size_t)parity * 4 + k) * nb + blockIdx.x, mine[k], gen);
  }
  if (blockIdx.x == 0) {
    for (unsigned idx = threadIdx.x; idx < NV * nb; idx += T) {
      const unsigned k = idx / nb, i = idx - k * nb;
      double val; unsigned long long g;
      do { ld_rec<2>(recs + ((size_t)parity * 4 + k) * nb + i, val, g); } while (g != gen);
      red[idx] = val;
    }
    __syncthreads();
    if (warp < NV) {
      double s = 0.0;
      for (unsigned i = lane; i < nb; i += 32) s += red[warp * nb + i];
      s = warp_sum(s);
      if (lane == 0) st_rec<2>(bc + parity * 4 + warp, s, gen);
    }
  }
  if (warp == 0 && lane < NV) {
    double val; unsigned long long g;
    do { ld_rec<2>(bc + parity * 4 + lane, val, g); } while (g != gen);
    smem[64 + lane] = val;
  }
  __syncthreads();
  for (int k = 0; k < NV; ++k) v[k] = smem[64 + k];
  __syncthreads();
}
// MODE 8: fence-free all-gather: every CTA polls all tagged records itself, one record per thread.
template <int NV> __device__ void allreduce_tagged_ag(double (&v)[NV], Rec* recs, unsigned long long gen, double* smem, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nb = gridDim.x;
  const int parity = (int)(gen & 1ull);
  for (int k = 0; k < NV; ++k) { const double s = warp_sum(v[k]); if (lane == 0) smem[k * W + warp] = s; }
  __syncthreads();
  if (warp == 0) {
    double mine[NV];
    for (int k = 0; k < NV; ++k) { double s = lane < W ? smem[k * W + lane] : 0.0; mine[k] = warp_sum(s); }
    if (lane == 0) for (int k = 0; k < NV; ++k) st_rec<2>(recs + ((size_t)parity * 4 + k) * nb + blockIdx.x, mine[k], gen);
  }
  for (unsigned idx = threadIdx.x; idx < NV * nb; idx += T) {
    const unsigned k = idx / nb, i = idx - k * nb;
    double val; unsigned long long g;
    do { ld_rec<2>(recs + ((size_t)parity * 4 + k) * nb + i, val, g); } while (g != gen);
    red[idx] = val;
  }
  __syncthreads();
  if (warp < NV) {
    double s = 0.0;
    for (unsigned i = lane; i < nb; i += 32) s += red[warp * nb + i];
    s = warp_sum(s);
    if (lane == 0) smem[64 + warp] = s;
  }
  __syncthreads();
  for (int k = 0; k < NV; ++k) v[k] = smem[64 + k];
  __syncthreads();
}
// MODE 3: cooperative groups grid.sync + partial array ; 4: atomic counter barrier (old kernel)
template <int MODE> __global__ void __launch_bounds__(T, 1) bench(Rec* recs, double* partials, unsigned* bar, unsigned long long gen0, int n, double* out) {
  __shared__ double smem[72];
  double v[3] = {1.0, 2.0, 3.0};
  unsigned long long gen = gen0;
  if (MODE <= 2) {
    for (int i = 0; i < n; ++i) { allreduce<MODE, 3>(v, recs, gen++, smem); v[0] = v[0] * 1e-3 + 1.0; }
  } else if (MODE == 8) {
    __shared__ double red[3 * 160];
    for (int i = 0; i < n; ++i) { allreduce_tagged_ag<3>(v, recs, gen++, smem, red); v[0] = v[0] * 1e-3 + 1.0; }
  } else if (MODE == 7) {
    __shared__ double red[3 * 160];
    for (int i = 0; i < n; ++i) { allreduce_tagged<3>(v, recs, recs + 8 * 160, gen++, smem, red); v[0] = v[0] * 1e-3 + 1.0; }
  } else if (MODE >= 5) {
    for (int i = 0; i < n; ++i) { allreduce_last<MODE, 3>(v, partials, bar, recs, gen++, smem); v[0] = v[0] * 1e-3 + 1.0; }
  } else if (MODE == 3) {
    cg::grid_group g = cg::this_grid();
    for (int i = 0; i < n; ++i) {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      for (int k = 0; k < 3; ++k) { const double s = warp_sum(v[k]); if (lane == 0) smem[k * W + warp] = s; }
      __syncthreads();
      if (warp == 0) for (int k = 0; k < 3; ++k) { double s = lane < W ? smem[k * W + lane] : 0.0; s = warp_sum(s); if (lane == 0) partials[((i & 1) * 4 + k) * gridDim.x + blockIdx.x] = s; }
      g.sync();
      if (warp < 3) { double s = 0; for (unsigned j = lane; j < gridDim.x; j += 32) s += __ldcg(partials + ((i & 1) * 4 + warp) * gridDim.x + j); s = warp_sum(s); if (lane == 0) smem[64 + warp] = s; }
      __syncthreads();
      for (int k = 0; k < 3; ++k) v[k] = smem[64 + k];
      __syncthreads();
      v[0] = v[0] * 1e-3 + 1.0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = v[0];
}
template <int MODE> float run(int nb, int n, Rec* recs, double* partials, unsigned* bar, double* out, unsigned long long& gen) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    void* args[] = {&recs, &partials, &bar, &gen, &n, &out};
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)bench<MODE>, dim3(nb), dim3(T), args, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return -1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
    gen += n + 8;
  }
  return best * 1e3f / n;
}
int main() {
  Rec* recs; double *partials, *out; unsigned* bar;
  cudaMalloc(&recs, sizeof(Rec) * 8 * 256); cudaMemset(recs, 0, sizeof(Rec) * 8 * 256);
  cudaMalloc(&partials, sizeof(double) * 8 * 256); cudaMalloc(&out, 8); cudaMalloc(&bar, 8); cudaMemset(bar, 0, 8);
  unsigned long long gen = 1;
  for (int nb : {114, 148}) {
    const int n = 2000;
    printf("nb=%d  volatile+threadfence: %.3f us | release/acquire.gpu: %.3f us | relaxed.gpu+fence.acq_rel: %.3f us | cg grid.sync: %.3f us\n", nb,
           run<0>(nb, n, recs, partials, bar, out, gen), run<1>(nb, n, recs, partials, bar, out, gen),
           run<2>(nb, n, recs, partials, bar, out, gen), run<3>(nb, n, recs, partials, bar, out, gen));
    printf("nb=%d  last-arriver atom.acq_rel: %.3f us | last-arriver fences+relaxed atomic: %.3f us\n", nb,
           run<5>(nb, n, recs, partials, bar, out, gen), run<6>(nb, n, recs, partials, bar, out, gen));
    printf("nb=%d  fence-free tagged (CTA 0 reduces): %.3f us | fence-free tagged all-gather: %.3f us\n", nb, run<7>(nb, n, recs, partials, bar, out, gen), run<8>(nb, n, recs, partials, bar, out, gen));
  }
  return 0;
}
