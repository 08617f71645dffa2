// C ABI glue (include/mono_abi.h): context life cycle, buffers, the split-step driver, measurement.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include "mono_ctx.h"

static thread_local std::string g_create_error;

int mono_fail(mono_ctx* c, int code, const std::string& msg) {
  if (c)
    c->err = msg;
  else
    g_create_error = msg;
  return code;
}

namespace {

__global__ void copy_kernel(int64_t n, const double* __restrict__ src, double* __restrict__ d0, double* __restrict__ d1,
                            double* __restrict__ d2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = src[i];
    if (d0) d0[i] = v;
    if (d1) d1[i] = v;
    if (d2) d2[i] = v;
  }
}

__global__ void dfma_bench_kernel(double* out, int iters) {
  // 16 independent FMA chains per thread: saturates the fp64 pipe regardless of its latency
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  const double m = 1.0000000001, c = 1e-12;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

int copy3(mono_ctx* c, int64_t n, const double* src, double* d0, double* d1, double* d2) {
  if (n <= 0) return MONO_OK;
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, (int64_t)c->n_sm * 8);
  copy_kernel<<<blocks, threads, 0, c->stream>>>(n, src, d0, d1, d2);
  c->launches++;
  MONO_CUDA(c, cudaGetLastError());
  return MONO_OK;
}

// After a fused split step the PDE solution x is authoritative; refresh the mirrors the reference keeps
// (states[v_index], v_ode, v_) before any un-fused primitive or read-back looks at them.
int canonicalize(mono_ctx* c) {
  if (!c->fused_pending) return MONO_OK;
  c->fused_pending = false;
  double* row = c->has_ode ? c->states + (int64_t)c->v_index * c->ld : nullptr;
  const int64_t n = c->has_ode ? std::min(c->npts, c->n_local) : c->n_local;
  return copy3(c, n, c->x, row, c->has_ode ? c->v_ode : nullptr, c->v_prev);
}

bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// Host -> device on the context's stream.  Pageable source: the call returns after the copy (self-contained).
// Page-locked source (mono_host_alloc): the copy is asynchronous - the caller must not modify the buffer before
// the next synchronising call on this context (mono_sync, any get/info call).
int h2d(mono_ctx* c, double* dst, const double* src, int64_t n) {
  MONO_CUDA(c, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (!is_pinned(src)) MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  return MONO_OK;
}

int d2h(mono_ctx* c, double* dst, const double* src, int64_t n) {
  MONO_CUDA(c, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  return MONO_OK;
}

int stage_begin(mono_ctx* c) {
  if (!c->stage_timing) return MONO_OK;
  while (c->ev_used + 2 > c->ev_pool.size()) {
    cudaEvent_t e;
    MONO_CUDA(c, cudaEventCreate(&e));
    c->ev_pool.push_back(e);
  }
  MONO_CUDA(c, cudaEventRecord(c->ev_pool[c->ev_used], c->stream));
  return MONO_OK;
}

int stage_end(mono_ctx* c, int which) {
  if (!c->stage_timing) return MONO_OK;
  MONO_CUDA(c, cudaEventRecord(c->ev_pool[c->ev_used + 1], c->stream));
  c->ev_used += 2;
  c->ev_tags.push_back(which);
  return MONO_OK;
}

int drain_stage_events(mono_ctx* c) {
  if (c->ev_used == 0) return MONO_OK;
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  for (size_t k = 0; k + 1 < c->ev_used; k += 2) {
    float ms = 0.f;
    MONO_CUDA(c, cudaEventElapsedTime(&ms, c->ev_pool[k], c->ev_pool[k + 1]));
    c->stage_ms[c->ev_tags[k / 2]] += ms;
  }
  c->ev_used = 0;
  c->ev_tags.clear();
  return MONO_OK;
}

}  // namespace

extern "C" {

int mono_abi_version(void) { return MONO_ABI_VERSION; }

const char* mono_last_error(const mono_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mono_ctx_create(int device, mono_ctx** out) {
  if (!out) return mono_fail(nullptr, MONO_E_INVALID, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return mono_fail(nullptr, MONO_E_CUDA,
                     std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return mono_fail(nullptr, MONO_E_INVALID, "device ordinal out of range");
  mono_ctx* c = new mono_ctx();
  c->device = device;
  auto bail = [&](const char* what, cudaError_t err) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(err);
    delete c;
    return mono_fail(nullptr, MONO_E_CUDA, m);
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  c->n_sm = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaMalloc(&c->ksp_dev, sizeof(KspResult))) != cudaSuccess) return bail("cudaMalloc", e);
  cudaMemset(c->ksp_dev, 0, sizeof(KspResult));
  if ((e = cudaMallocHost(&c->ksp_host, sizeof(KspResult))) != cudaSuccess) return bail("cudaMallocHost", e);
  memset(c->ksp_host, 0, sizeof(KspResult));
  *out = c;
  return MONO_OK;
}

int mono_ctx_destroy(mono_ctx* c) {
  if (!c) return MONO_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (void* p : {(void*)c->states, (void*)c->v_ode, (void*)c->params_dev, (void*)c->slice_ptr, (void*)c->cols,
                  (void*)c->mass, (void*)c->stiff, (void*)c->A, (void*)c->B, (void*)c->dinv, (void*)c->x,
                  (void*)c->v_prev, (void*)c->work[0], (void*)c->work[1], (void*)c->work[2], (void*)c->work[3],
                  (void*)c->work[4], (void*)c->work[5], (void*)c->work[6], (void*)c->work[7], (void*)c->work[8], (void*)c->work[9],
                  (void*)c->stim_vec,
                  (void*)c->gen_state, (void*)c->send_of_row_dev, (void*)c->send_ents_dev, (void*)c->slice_send_dev, (void*)c->region_table, (void*)c->region_of_node,
                  (void*)c->recs, (void*)c->timeline_dev,
                  (void*)c->ksp_dev, (void*)c->probes_dev,
                  (void*)c->probe_vals_dev, (void*)c->probe_act_dev, (void*)c->flush_buf, (void*)c->send_idx_dev,
                  (void*)c->send_buf, (void*)c->red_buf, (void*)c->pat_dev, (void*)c->dict_off_dev, (void*)c->dict_w_dev,
                  (void*)c->dict_src_dev, (void*)c->dict_A_dev, (void*)c->dict_B_dev, (void*)c->dict_cl_dev, (void*)c->nd_rows_dev,
                  (void*)c->nd_w_dev, (void*)c->nd_cols_dev, (void*)c->nd_src_dev, (void*)c->nd_A_dev, (void*)c->nd_B_dev,
                  (void*)c->nd_cta_ptr_dev, (void*)c->dict_rep_dev, (void*)c->dict_dinv_dev, (void*)c->slice_stim_dev, (void*)c->actmap_dev, (void*)c->minmax_dev, (void*)c->snap_dev})
    if (p) cudaFree(p);
  for (void* p : c->stim_allocs) cudaFree(p);
  if (c->ksp_host) cudaFreeHost(c->ksp_host);
  for (auto& t : c->timers)
    for (auto& ev : t)
      if (ev) cudaEventDestroy(ev);
  for (auto ev : c->ev_pool) cudaEventDestroy(ev);
  for (auto ev : c->marks)
    if (ev) cudaEventDestroy(ev);

  halo_destroy(c);  // closes the peer mappings; must precede freeing our own exported allocation
  if (c->exch) cudaFree(c->exch);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return MONO_OK;
}

int mono_host_alloc(int64_t nbytes, void** out) {
  if (!out || nbytes < 0) return mono_fail(nullptr, MONO_E_INVALID, "bad arguments");
  *out = nullptr;
  cudaError_t e = cudaMallocHost(out, (size_t)std::max<int64_t>(nbytes, 8));
  if (e != cudaSuccess) return mono_fail(nullptr, MONO_E_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
  return MONO_OK;
}

int mono_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
  return MONO_OK;
}

int mono_sync(mono_ctx* c) {
  MONO_CHECK(c, c != nullptr, "ctx is NULL");
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  return MONO_OK;
}

int mono_device_info(mono_ctx* c, int* n_sm, int* cc_major, int* cc_minor, int64_t* mem_bytes) {
  MONO_NEED_CTX(c);
  cudaDeviceProp prop;
  MONO_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
  if (n_sm) *n_sm = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (mem_bytes) *mem_bytes = (int64_t)prop.totalGlobalMem;
  return MONO_OK;
}

// ---------------------------------------------------------------------------------------------- ODE
int mono_ode_create(mono_ctx* c, int model_id, int scheme_id, int64_t num_points, int v_index) {
  MONO_CHECK(c, !c->has_ode, "ODE stage already created");
  int ns = 0, np = 0;
  MONO_CHECK(c, ode_model_dims(model_id, &ns, &np) == 0, "unknown model id");
  MONO_CHECK(c, scheme_id == MONO_SCHEME_FORWARD_EULER || scheme_id == MONO_SCHEME_GRL1, "unknown scheme id");
  MONO_CHECK(c, num_points >= 0, "negative num_points");
  MONO_CHECK(c, v_index >= 0 && v_index < ns, "v_index out of range");
  MONO_CHECK(c, !c->has_pde || num_points == c->n_local, "num_points must equal owned+ghost dofs of the PDE stage");
  MONO_CUDA(c, cudaSetDevice(c->device));
  c->model_id = model_id;
  c->scheme_id = scheme_id;
  c->ns = ns;
  c->np = np;
  c->nd = ode_model_num_derived(model_id, scheme_id);
  c->v_index = v_index;
  c->npts = num_points;
  c->ld = ((num_points + 31) / 32) * 32;  // 256-byte aligned rows
  const int64_t ld = std::max<int64_t>(c->ld, 32);
  MONO_CUDA(c, cudaMalloc(&c->states, sizeof(double) * ns * ld));
  MONO_CUDA(c, cudaMemsetAsync(c->states, 0, sizeof(double) * ns * ld, c->stream));
  MONO_CUDA(c, cudaMalloc(&c->v_ode, sizeof(double) * ld));
  MONO_CUDA(c, cudaMemsetAsync(c->v_ode, 0, sizeof(double) * ld, c->stream));
  c->has_ode = true;
  return MONO_OK;
}

int mono_ode_set_states(mono_ctx* c, const double* states, int64_t ld) {
  MONO_CHECK(c, c->has_ode, "no ODE stage");
  MONO_CHECK(c, ld >= c->npts, "ld < num_points");
  int rc = canonicalize(c);
  if (rc) return rc;
  if (c->npts == 0) return MONO_OK;
  MONO_CUDA(c, cudaMemcpy2DAsync(c->states, c->ld * sizeof(double), states, ld * sizeof(double), c->npts * sizeof(double),
                                 c->ns, cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  return MONO_OK;
}

int mono_ode_get_states(mono_ctx* c, double* states, int64_t ld) {
  MONO_CHECK(c, c->has_ode, "no ODE stage");
  MONO_CHECK(c, ld >= c->npts, "ld < num_points");
  int rc = canonicalize(c);
  if (rc) return rc;
  if (c->npts == 0) return MONO_OK;
  MONO_CUDA(c, cudaMemcpy2DAsync(states, ld * sizeof(double), c->states, c->ld * sizeof(double), c->npts * sizeof(double),
                                 c->ns, cudaMemcpyDeviceToHost, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  return MONO_OK;
}

int mono_ode_set_state_row(mono_ctx* c, int row, const double* values) {
  MONO_CHECK(c, c->has_ode && row >= 0 && row < c->ns, "bad state row");
  int rc = canonicalize(c);
  if (rc) return rc;
  return h2d(c, c->states + (int64_t)row * c->ld, values, c->npts);
}

int mono_ode_get_state_row(mono_ctx* c, int row, double* values) {
  MONO_CHECK(c, c->has_ode && row >= 0 && row < c->ns, "bad state row");
  int rc = canonicalize(c);
  if (rc) return rc;
  return d2h(c, values, c->states + (int64_t)row * c->ld, c->npts);
}

int mono_ode_set_params(mono_ctx* c, const double* params, int num_params, int per_node, int64_t ld, const double* derived,
                        int n_derived) {
  MONO_CHECK(c, c->has_ode, "no ODE stage");
  MONO_CHECK(c, num_params == c->np, "wrong number of parameters for this model");
  if (c->region_of_node) {  // back from per-region parameters
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->region_of_node);
    c->region_of_node = nullptr;
  }
  if (per_node) {
    MONO_CHECK(c, ld >= c->npts, "ld < num_points");
    if (!c->params_dev) MONO_CUDA(c, cudaMalloc(&c->params_dev, sizeof(double) * c->np * std::max<int64_t>(c->ld, 32)));
    if (c->npts > 0) {
      MONO_CUDA(c, cudaMemcpy2DAsync(c->params_dev, c->ld * sizeof(double), params, ld * sizeof(double),
                                     c->npts * sizeof(double), c->np, cudaMemcpyHostToDevice, c->stream));
      MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    c->per_node = true;
  } else {
    MONO_CHECK(c, n_derived <= c->nd, "too many derived constants");
    MONO_CHECK(c, n_derived == 0 || derived != nullptr, "derived is NULL");
    c->params_host.assign(c->np + c->nd, 0.0);
    for (int k = 0; k < c->np; ++k) c->params_host[k] = params[k];
    for (int k = 0; k < n_derived; ++k) c->params_host[c->np + k] = derived[k];
    c->per_node = false;
  }
  c->have_params = true;
  return MONO_OK;
}

int mono_ode_set_region_params(mono_ctx* c, int n_regions, const double* params, int num_params, const double* derived, int n_derived,
                               const int32_t* region_of_node) {
  MONO_CHECK(c, c->has_ode, "no ODE stage");
  MONO_CHECK(c, n_regions >= 1 && num_params == c->np, "wrong number of regions / parameters for this model");
  MONO_CHECK(c, n_derived <= c->nd && (n_derived == 0 || derived != nullptr), "bad derived constants");
  const int stride = c->np + std::max(c->nd, 1);
  std::vector<double> tbl((size_t)n_regions * stride, 0.0);
  for (int r = 0; r < n_regions; ++r) {
    for (int k = 0; k < c->np; ++k) tbl[(size_t)r * stride + k] = params[(size_t)r * c->np + k];
    for (int k = 0; k < n_derived; ++k) tbl[(size_t)r * stride + c->np + k] = derived[(size_t)r * n_derived + k];
  }
  if (region_of_node != nullptr) {  // (NULL: keep the node -> region map, only the parameter values changed)
    for (int64_t i = 0; i < c->npts; ++i) MONO_CHECK(c, region_of_node[i] >= 0 && region_of_node[i] < n_regions, "region index out of range");
    if (!c->region_of_node) MONO_CUDA(c, cudaMalloc(&c->region_of_node, sizeof(int32_t) * std::max<int64_t>(c->ld, 32)));
    MONO_CUDA(c, cudaMemcpyAsync(c->region_of_node, region_of_node, sizeof(int32_t) * c->npts, cudaMemcpyHostToDevice, c->stream));
  } else {
    MONO_CHECK(c, c->region_of_node != nullptr && n_regions == c->n_regions, "no region map set yet");
  }
  if (c->region_table && n_regions != c->n_regions) {
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->region_table);
    c->region_table = nullptr;
  }
  if (!c->region_table) MONO_CUDA(c, cudaMalloc(&c->region_table, sizeof(double) * tbl.size()));
  MONO_CUDA(c, cudaMemcpyAsync(c->region_table, tbl.data(), sizeof(double) * tbl.size(), cudaMemcpyHostToDevice, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_regions = n_regions;
  c->per_node = false;
  c->have_params = true;
  return MONO_OK;
}

int mono_ode_step(mono_ctx* c, double t0, double dt) {
  MONO_CHECK(c, c->has_ode && c->have_params, "ODE stage not initialised (create + set_params)");
  int rc = canonicalize(c);
  if (rc) return rc;
  return ode_launch(c, t0, dt, nullptr, nullptr, nullptr);
}

int mono_ode_to_dolfin(mono_ctx* c) {
  MONO_CHECK(c, c->has_ode, "no ODE stage");
  int rc = canonicalize(c);
  if (rc) return rc;
  return copy3(c, c->npts, c->states + (int64_t)c->v_index * c->ld, c->v_ode, nullptr, nullptr);
}

int mono_ode_from_dolfin(mono_ctx* c) {
  MONO_CHECK(c, c->has_ode, "no ODE stage");
  int rc = canonicalize(c);
  if (rc) return rc;
  return copy3(c, c->npts, c->v_ode, c->states + (int64_t)c->v_index * c->ld, nullptr, nullptr);
}

int mono_ode_to_pde(mono_ctx* c) {
  MONO_CHECK(c, c->has_ode && c->has_pde, "needs ODE and PDE stages");
  int rc = canonicalize(c);
  if (rc) return rc;
  return copy3(c, c->n_local, c->v_ode, c->x, nullptr, nullptr);
}

int mono_pde_to_ode(mono_ctx* c) {
  MONO_CHECK(c, c->has_ode && c->has_pde, "needs ODE and PDE stages");
  int rc = canonicalize(c);
  if (rc) return rc;
  return copy3(c, c->n_local, c->x, c->v_ode, nullptr, nullptr);
}

int mono_get_v_ode(mono_ctx* c, double* v) {
  MONO_CHECK(c, c->has_ode, "no ODE stage");
  int rc = canonicalize(c);
  if (rc) return rc;
  return d2h(c, v, c->v_ode, c->npts);
}

int mono_set_v_ode(mono_ctx* c, const double* v) {
  MONO_CHECK(c, c->has_ode, "no ODE stage");
  int rc = canonicalize(c);
  if (rc) return rc;
  return h2d(c, c->v_ode, v, c->npts);
}

// ---------------------------------------------------------------------------------------------- PDE
int mono_pde_set_matrices(mono_ctx* c, int64_t n_owned, int64_t n_ghost, const int64_t* indptr, const int32_t* indices,
                          const double* mass, const double* stiff) {
  MONO_CHECK(c, !c->has_pde, "PDE matrices already set");
  MONO_CHECK(c, n_owned >= 0 && n_ghost >= 0, "negative sizes");
  MONO_CHECK(c, n_owned + n_ghost < (int64_t)std::numeric_limits<int32_t>::max(), "local dof count exceeds int32");
  MONO_CHECK(c, !c->has_ode || c->npts == n_owned + n_ghost, "owned+ghost must equal the ODE stage's num_points");
  MONO_CUDA(c, cudaSetDevice(c->device));
  c->n_owned = n_owned;
  c->n_ghost = n_ghost;
  c->n_local = n_owned + n_ghost;
  int rc = pde_build_sell(c, indptr, indices, mass, stiff);
  if (rc) return rc;
  const int64_t nl = std::max<int64_t>(c->n_local, 32);
  for (double** p : {&c->x, &c->v_prev, &c->dinv}) {
    // (+2: the ring kernel's bulk copies are 16-byte granular and may read one element past an odd-sized vector)
    MONO_CUDA(c, cudaMalloc(p, sizeof(double) * (nl + 2)));
    MONO_CUDA(c, cudaMemsetAsync(*p, 0, sizeof(double) * (nl + 2), c->stream));
  }
  c->has_pde = true;
  rc = pde_setup_launch_config(c);
  return rc;
}

int mono_pde_config(mono_ctx* c, double C_m, double theta, double rtol, double atol, int max_it, int pc_type, int norm_type,
                    int x0_mode) {
  MONO_CHECK(c, theta >= 0.0 && theta <= 1.0, "theta outside [0,1]");
  MONO_CHECK(c, pc_type == MONO_PC_NONE || pc_type == MONO_PC_JACOBI || pc_type == MONO_PC_CHEBYSHEV, "unknown pc_type");
  MONO_CHECK(c, norm_type >= 0 && norm_type <= 2, "unknown norm_type");
  MONO_CHECK(c, x0_mode == MONO_X0_ZERO || x0_mode == MONO_X0_PREVIOUS, "unknown x0_mode");
  MONO_CHECK(c, max_it >= 0, "negative max_it");
  c->C_m = C_m;
  c->theta = theta;
  c->rtol = rtol;
  c->atol = atol;
  c->max_it = max_it;
  c->pc_type = pc_type;
  c->norm_type = norm_type;
  c->x0_mode = x0_mode;
  c->mode_dirty = true;
  c->have_dt = false;  // matrices depend on C_m / theta / pc
  if (c->has_pde) {
    const int nb = pc_type == MONO_PC_CHEBYSHEV ? 3 * c->cheb_k : 2;
    if (nb != c->exch_nb) {  // the polynomial preconditioner exchanges one vector per step: more tagged buffers
      MONO_CHECK(c, !c->peers_ready, "choose the preconditioner before mono_set_halo (peers have mapped the exchange buffers)");
      return pde_setup_launch_config(c);
    }
  }
  return MONO_OK;
}

int mono_pde_set_chebyshev(mono_ctx* c, int steps, double kappa) {
  MONO_CHECK(c, steps >= 1 && steps <= kMaxCheb, "chebyshev steps must be 1..4");
  MONO_CHECK(c, kappa > 1.0, "kappa must exceed 1");
  const bool resize = steps != c->cheb_k && c->pc_type == MONO_PC_CHEBYSHEV && c->has_pde;
  MONO_CHECK(c, !(resize && c->peers_ready), "choose the preconditioner before mono_set_halo");
  c->cheb_k = steps;
  c->cheb_kappa = kappa;
  c->mode_dirty = true;
  c->have_dt = false;
  return resize ? pde_setup_launch_config(c) : MONO_OK;
}

int mono_pde_set_ksp_type(mono_ctx* c, int ksp_type) {
  MONO_CHECK(c, ksp_type == MONO_KSP_CG || ksp_type == MONO_KSP_PIPECG, "unknown ksp_type");
  c->ksp_type = ksp_type;
  c->mode_dirty = true;
  return MONO_OK;
}

int mono_pde_set_dt(mono_ctx* c, double dt) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  return pde_update_matrices(c, dt);
}

int mono_stim_add(mono_ctx* c, int64_t nnz, const int32_t* idx, const double* val, double t_start, double t_end,
                  double amplitude) {
  MONO_CHECK(c, c->has_pde, "set matrices before adding stimuli");
  MONO_CHECK(c, nnz >= 0, "negative nnz");
  MONO_CHECK(c, nnz == 0 || (idx && val), "stimulus arrays are NULL");
  for (int64_t k = 0; k < nnz; ++k) MONO_CHECK(c, idx[k] >= 0 && idx[k] < c->n_owned, "stimulus index is not an owned dof");
  StimDev s{};
  s.nnz = nnz;
  s.t_start = t_start;
  s.t_end = t_end;
  s.amp = amplitude;
  if (nnz > 0) {
    int32_t* di = nullptr;
    double* dv = nullptr;
    MONO_CUDA(c, cudaMalloc(&di, sizeof(int32_t) * nnz));
    MONO_CUDA(c, cudaMalloc(&dv, sizeof(double) * nnz));
    c->stim_allocs.push_back(di);
    c->stim_allocs.push_back(dv);
    MONO_CUDA(c, cudaMemcpyAsync(di, idx, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaMemcpyAsync(dv, val, sizeof(double) * nnz, cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    s.idx = di;
    s.val = dv;
  }
  c->stims_host.push_back(s);
  return (int)c->stims_host.size() - 1;
}

int mono_stim_set_amplitude(mono_ctx* c, int id, double amplitude) {
  MONO_CHECK(c, id >= 0 && id < (int)c->stims_host.size(), "bad stimulus id");
  c->stims_host[id].amp = amplitude;  // pde_launch_step notices the change and rebuilds the dense source vector
  return MONO_OK;
}

int mono_stim_set_window(mono_ctx* c, int id, double t_start, double t_end) {
  MONO_CHECK(c, id >= 0 && id < (int)c->stims_host.size(), "bad stimulus id");
  c->stims_host[id].t_start = t_start;
  c->stims_host[id].t_end = t_end;
  return MONO_OK;
}

static int pde_step_impl(mono_ctx* c, double t0, double t1) {
  const double dt = t1 - t0;
  const double t = t0 + c->theta * dt;  // base_model.py:216-223: the stimulus is evaluated at the theta point
  if (!c->have_dt || std::fabs(dt - c->cur_dt) >= 1e-12) {  // base_model.py:225-230
    int rc = pde_update_matrices(c, dt);
    if (rc) return rc;
  }
  // the kernel also refreshes the ghosts of x (state.x.scatter_forward(), base_model.py:242) before it ends
  return pde_launch_step(c, t, dt);
}

int mono_pde_step(mono_ctx* c, double t0, double t1) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  int rc = canonicalize(c);
  if (rc) return rc;
  return pde_step_impl(c, t0, t1);
}

int mono_pde_assign_previous(mono_ctx* c) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  int rc = canonicalize(c);
  if (rc) return rc;
  return copy3(c, c->n_local, c->x, c->v_prev, nullptr, nullptr);
}

int mono_get_v(mono_ctx* c, double* v) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  return d2h(c, v, c->x, c->n_local);
}

int mono_set_v(mono_ctx* c, const double* v) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  int rc = canonicalize(c);
  if (rc) return rc;
  return h2d(c, c->x, v, c->n_local);
}

int mono_get_v_prev(mono_ctx* c, double* v) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  int rc = canonicalize(c);
  if (rc) return rc;
  return d2h(c, v, c->v_prev, c->n_local);
}

int mono_set_v_prev(mono_ctx* c, const double* v) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  int rc = canonicalize(c);
  if (rc) return rc;
  return h2d(c, c->v_prev, v, c->n_local);
}

int mono_ksp_info(mono_ctx* c, int* iterations, double* residual_norm, int* reason) {
  MONO_CUDA(c, cudaMemcpyAsync(c->ksp_host, c->ksp_dev, sizeof(KspResult), cudaMemcpyDeviceToHost, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  if (iterations) *iterations = c->ksp_host->iterations;
  if (residual_norm) *residual_norm = c->ksp_host->rnorm;
  if (reason) *reason = c->ksp_host->reason;
  if (c->ksp_host->error) return mono_fail(c, MONO_E_CUDA, "PDE kernel: grid synchronisation timed out");
  return MONO_OK;
}

int mono_ksp_total_iterations(mono_ctx* c, int64_t* total, int64_t* solves) {
  MONO_CUDA(c, cudaMemcpyAsync(c->ksp_host, c->ksp_dev, sizeof(KspResult), cudaMemcpyDeviceToHost, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  if (total) *total = c->ksp_host->total_iterations;
  if (solves) *solves = c->ksp_host->solves;
  return MONO_OK;
}

int mono_pde_dictionary_info(mono_ctx* c, int* n_patterns, double* rows_covered, int* active) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  if (n_patterns) *n_patterns = c->n_pat;
  if (rows_covered) *rows_covered = c->dict_cover;
  if (active) *active = c->dict_active ? 1 : 0;
  return MONO_OK;
}

// -------------------------------------------------------------------------------------------- fused
static int split_step_impl(mono_ctx* c, double t0, double t1, double theta) {
  const double dt = t1 - t0;
  // ode.step(t0, theta*dt); to_dolfin; ode_to_pde; assign_previous   (monodomain_solver.py:66-79)
  // -> K1 reads V from the authoritative buffer and writes the new V straight into v_.
  const double* v_in = c->fused_pending ? c->x : nullptr;
  int rc = stage_begin(c);
  if (rc) return rc;
  rc = ode_launch(c, t0, theta * dt, v_in, c->v_prev, nullptr);
  if (rc) return rc;
  rc = stage_end(c, 0);
  if (rc) return rc;
  // pde.step((t0, t1))                                                 (monodomain_solver.py:84)
  rc = stage_begin(c);
  if (rc) return rc;
  rc = pde_step_impl(c, t0, t1);
  if (rc) return rc;
  rc = stage_end(c, 1);
  if (rc) return rc;
  // pde_to_ode; from_dolfin; assign_previous (:88-97) are fused away: x is now authoritative.
  c->fused_pending = true;
  if (std::fabs(theta - 1.0) > 1e-8) {
    // corrective ODE step (Strang), :98-113: ode.step(t0+theta*dt, (1-theta)*dt); new V -> v_ode, v, v_
    rc = stage_begin(c);
    if (rc) return rc;
    rc = ode_launch(c, t0 + theta * dt, (1.0 - theta) * dt, c->x, c->x, nullptr);
    if (rc) return rc;
    rc = stage_end(c, 0);
    if (rc) return rc;
  }
  if (!c->probes_host.empty() || c->actmap_enabled || c->minmax_enabled) rc = probes_launch(c, t0);
  if (c->stage_timing) c->stage_steps++;
  if (c->stage_timing && c->ev_used >= 4096) rc = drain_stage_events(c);
  return rc;
}

int mono_split_step(mono_ctx* c, double t0, double t1, double theta_split) {
  MONO_CHECK(c, c->has_ode && c->has_pde && c->have_params, "split step needs ODE stage, parameters and PDE matrices");
  MONO_CHECK(c, c->npts == c->n_local, "ODE and PDE stages disagree on the number of dofs");
  return split_step_impl(c, t0, t1, theta_split);
}

int mono_split_solve(mono_ctx* c, double t0, double dt, int64_t nsteps, double theta_split) {
  MONO_CHECK(c, c->has_ode && c->has_pde && c->have_params, "split solve needs ODE stage, parameters and PDE matrices");
  MONO_CHECK(c, c->npts == c->n_local, "ODE and PDE stages disagree on the number of dofs");
  double t = t0;
  for (int64_t k = 0; k < nsteps; ++k) {
    const double tn = t + dt;  // same accumulation as the reference loop (monodomain_solver.py:44-51)
    int rc = split_step_impl(c, t, tn, theta_split);
    if (rc) return rc;
    t = tn;
  }
  return MONO_OK;
}

// ---------------------------------------------------------------------------------------- observers
int mono_probe_add(mono_ctx* c, int n_nodes, const int32_t* nodes, const double* weights) {
  MONO_CHECK(c, c->has_pde, "set matrices before adding probes");
  MONO_CHECK(c, n_nodes >= 1 && n_nodes <= 4 && nodes && weights, "a probe has 1..4 nodes");
  ProbeDev p{};
  p.n = n_nodes;
  for (int k = 0; k < n_nodes; ++k) {
    MONO_CHECK(c, nodes[k] >= 0 && nodes[k] < c->n_local, "probe node out of range");
    p.node[k] = nodes[k];
    p.w[k] = weights[k];
  }
  c->probes_host.push_back(p);
  c->probes_dirty = true;
  return (int)c->probes_host.size() - 1;
}

int mono_probe_values(mono_ctx* c, double* values) {
  MONO_NEED_CTX(c);
  const int n = (int)c->probes_host.size();
  if (n == 0) return MONO_OK;
  int rc = probes_launch(c, -1.0e300);  // evaluates; cannot activate (t0 only stored when crossing)
  if (rc) return rc;
  return d2h(c, values, c->probe_vals_dev, n);
}

int mono_probe_activation(mono_ctx* c, double threshold) {
  MONO_NEED_CTX(c);
  c->act_enabled = true;
  c->act_threshold = threshold;
  return MONO_OK;
}

int mono_probe_activation_times(mono_ctx* c, double* times) {
  MONO_NEED_CTX(c);
  const int n = (int)c->probes_host.size();
  if (n == 0) return MONO_OK;
  if (c->probes_dirty) {
    for (int k = 0; k < n; ++k) times[k] = -1.0;
    return MONO_OK;
  }
  return d2h(c, times, c->probe_act_dev, n);
}

int mono_observe_config(mono_ctx* c, int activation_map, double threshold, int minmax) {
  MONO_CHECK(c, c->has_pde, "set matrices before configuring the observers");
  if (activation_map && !c->actmap_dev) {
    const int64_t n = std::max<int64_t>(c->n_owned, 1);
    MONO_CUDA(c, cudaMalloc(&c->actmap_dev, sizeof(double) * n));
    std::vector<double> neg((size_t)n, -1.0);
    MONO_CUDA(c, cudaMemcpyAsync(c->actmap_dev, neg.data(), sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  if (minmax && !c->minmax_dev) {
    MONO_CUDA(c, cudaMalloc(&c->minmax_dev, 2 * sizeof(unsigned long long)));
    const unsigned long long init[2] = {~0ull, 0ull};
    MONO_CUDA(c, cudaMemcpyAsync(c->minmax_dev, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  c->actmap_enabled = activation_map != 0;
  c->actmap_threshold = threshold;
  c->minmax_enabled = minmax != 0;
  return MONO_OK;
}

int mono_activation_map(mono_ctx* c, double* times_owned) {
  MONO_CHECK(c, c->has_pde && c->actmap_dev != nullptr, "activation map not enabled (mono_observe_config)");
  MONO_CHECK(c, times_owned != nullptr || c->n_owned == 0, "times_owned is NULL");
  if (c->n_owned == 0) return MONO_OK;
  return d2h(c, times_owned, c->actmap_dev, c->n_owned);
}

int mono_v_minmax(mono_ctx* c, double* vmin, double* vmax) {
  MONO_CHECK(c, c->has_pde && c->minmax_dev != nullptr, "min/max tracking not enabled (mono_observe_config)");
  unsigned long long b[2];
  MONO_CUDA(c, cudaMemcpyAsync(b, c->minmax_dev, sizeof(b), cudaMemcpyDeviceToHost, c->stream));
  MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  auto back = [](unsigned long long u) {
    const unsigned long long bits = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double v;
    memcpy(&v, &bits, sizeof(v));
    return v;
  };
  if (vmin) *vmin = b[0] == ~0ull ? std::numeric_limits<double>::infinity() : back(b[0]);
  if (vmax) *vmax = b[1] == 0ull ? -std::numeric_limits<double>::infinity() : back(b[1]);
  return MONO_OK;
}

int mono_get_v_strided(mono_ctx* c, int64_t offset, int64_t stride, int64_t count, double* out) {
  MONO_CHECK(c, c->has_pde, "no PDE matrices");
  MONO_CHECK(c, offset >= 0 && stride >= 1 && count >= 0, "bad offset / stride / count");
  MONO_CHECK(c, count == 0 || offset + (count - 1) * stride < c->n_local, "strided range exceeds the local dofs");
  MONO_CHECK(c, count == 0 || out != nullptr, "out is NULL");
  if (count == 0) return MONO_OK;
  if (count > c->snap_cap) {
    if (c->snap_dev) {
      MONO_CUDA(c, cudaStreamSynchronize(c->stream));
      cudaFree(c->snap_dev);
      c->snap_dev = nullptr;
    }
    MONO_CUDA(c, cudaMalloc(&c->snap_dev, sizeof(double) * count));
    c->snap_cap = count;
  }
  int rc = strided_pack_launch(c, offset, stride, count, c->snap_dev);
  if (rc) return rc;
  return d2h(c, out, c->snap_dev, count);
}

// -------------------------------------------------------------------------------------- measurement
int mono_timer_start(mono_ctx* c, int slot) {
  MONO_CHECK(c, slot >= 0 && slot < 8, "timer slot out of range");
  for (int k = 0; k < 2; ++k)
    if (!c->timers[slot][k]) MONO_CUDA(c, cudaEventCreate(&c->timers[slot][k]));
  MONO_CUDA(c, cudaEventRecord(c->timers[slot][0], c->stream));
  return MONO_OK;
}

int mono_timer_stop(mono_ctx* c, int slot) {
  MONO_CHECK(c, slot >= 0 && slot < 8 && c->timers[slot][1], "timer not started");
  MONO_CUDA(c, cudaEventRecord(c->timers[slot][1], c->stream));
  return MONO_OK;
}

int mono_timer_elapsed_ms(mono_ctx* c, int slot, float* ms) {
  MONO_CHECK(c, slot >= 0 && slot < 8 && c->timers[slot][1], "timer not started");
  MONO_CUDA(c, cudaEventSynchronize(c->timers[slot][1]));
  MONO_CUDA(c, cudaEventElapsedTime(ms, c->timers[slot][0], c->timers[slot][1]));
  return MONO_OK;
}

int mono_event_record(mono_ctx* c, int idx) {
  MONO_CHECK(c, idx >= 0 && idx < (1 << 24), "event index out of range");
  if ((size_t)idx >= c->marks.size()) c->marks.resize((size_t)idx + 1, nullptr);
  if (!c->marks[idx]) MONO_CUDA(c, cudaEventCreate(&c->marks[idx]));
  MONO_CUDA(c, cudaEventRecord(c->marks[idx], c->stream));
  return MONO_OK;
}

int mono_event_elapsed_ms(mono_ctx* c, int idx0, int idx1, float* ms) {
  MONO_CHECK(c, idx0 >= 0 && idx1 >= 0 && (size_t)idx0 < c->marks.size() && (size_t)idx1 < c->marks.size() &&
                    c->marks[idx0] && c->marks[idx1],
             "event mark not recorded");
  MONO_CUDA(c, cudaEventSynchronize(c->marks[idx1]));
  MONO_CUDA(c, cudaEventElapsedTime(ms, c->marks[idx0], c->marks[idx1]));
  return MONO_OK;
}

int mono_l2_flush(mono_ctx* c) {
  MONO_NEED_CTX(c);
  if (!c->flush_buf) {
    c->flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    MONO_CUDA(c, cudaMalloc(&c->flush_buf, c->flush_bytes));
  }
  MONO_CUDA(c, cudaMemsetAsync(c->flush_buf, 0x5a, c->flush_bytes, c->stream));
  return MONO_OK;
}

int mono_stage_timing(mono_ctx* c, int enable) {
  MONO_NEED_CTX(c);
  c->stage_timing = enable != 0;
  return MONO_OK;
}

int mono_stage_times_ms(mono_ctx* c, double* ms2, int64_t* steps, int reset) {
  MONO_NEED_CTX(c);
  int rc = drain_stage_events(c);
  if (rc) return rc;
  if (ms2) {
    ms2[0] = c->stage_ms[0];
    ms2[1] = c->stage_ms[1];
  }
  if (steps) *steps = c->stage_steps;
  if (reset) {
    c->stage_ms[0] = c->stage_ms[1] = 0.0;
    c->stage_steps = 0;
  }
  return MONO_OK;
}

int mono_bench_dfma(mono_ctx* c, double* tflops) {
  MONO_NEED_CTX(c);
  double* out = nullptr;
  MONO_CUDA(c, cudaMalloc(&out, sizeof(double)));
  const int blocks = c->n_sm * 8, threads = 256, iters = 4096;
  cudaEvent_t e0, e1;
  MONO_CUDA(c, cudaEventCreate(&e0));
  MONO_CUDA(c, cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    MONO_CUDA(c, cudaEventRecord(e0, c->stream));
    dfma_bench_kernel<<<blocks, threads, 0, c->stream>>>(out, iters);
    MONO_CUDA(c, cudaEventRecord(e1, c->stream));
    MONO_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    MONO_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 16.0 * iters * (double)blocks * threads;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (tflops) *tflops = best;
  return MONO_OK;
}

int mono_debug_timeline(mono_ctx* c, int enable, uint64_t* stamps64) {
  MONO_NEED_CTX(c);
  if (enable && !c->timeline_dev) {
    MONO_CUDA(c, cudaMalloc(&c->timeline_dev, sizeof(unsigned long long) * 64));
    MONO_CUDA(c, cudaMemsetAsync(c->timeline_dev, 0, sizeof(unsigned long long) * 64, c->stream));
  }
  if (stamps64 && c->timeline_dev) {
    MONO_CUDA(c, cudaMemcpyAsync(stamps64, c->timeline_dev, sizeof(unsigned long long) * 64, cudaMemcpyDeviceToHost, c->stream));
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  if (!enable && c->timeline_dev) {
    MONO_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->timeline_dev);
    c->timeline_dev = nullptr;
  }
  return MONO_OK;
}

int mono_bench_grid_sync(mono_ctx* c, int n, float* us_per_sync) {
  MONO_CHECK(c, n > 0, "n must be positive");
  return pde_bench_sync(c, n, us_per_sync);
}

int mono_launch_count(mono_ctx* c, int64_t* launches) {
  MONO_NEED_CTX(c);
  if (launches) *launches = c->launches;
  return MONO_OK;
}

}  // extern "C"
