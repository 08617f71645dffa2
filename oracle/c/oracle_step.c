/* ORACLE (test infrastructure, never a product path): the reference's operator-split monodomain step
 * restated in plain C with OpenMP, used (a) as a second, independently compiled checker of the NumPy
 * oracle and (b) as the multi-core CPU baseline that bench.py times (`cpu_baseline`, `--impl reference`).
 *
 * Follows, under /root/reference:
 *   src/beat/monodomain_solver.py:53-116   splitting order (Godunov theta=1 / Strang theta!=1)
 *   src/beat/odesolver.py:67-79            states[:] = fun(states, t, parameters, dt) over SoA (ns, N)
 *   src/beat/odesolver.py:164-170          V row <-> v_ode copies ; src/beat/utils.py:52-54 identity projection
 *   src/beat/monodomain_model.py:59-60     assign_previous ; :83-96 theta-rule form
 *   src/beat/base_model.py:208-245         time at theta point, RHS, KSP solve (restated as PETSc-style KSPCG
 *                                          with Jacobi: zero initial guess, preconditioned-norm test)
 * The cell-model scalar functions come from the generated oracle/c/<model>.c (oracle/gen_models.py).
 * No reference test covers the generated cell models; they are pinned on the published Niederer activation times
 * (demos/niederer_benchmark.py:315-325, reproduced within one dt: tests/test_oracle_niederer.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*step_fn)(const double *, double *, double, double, const double *);
void fhn_fe(const double *, double *, double, double, const double *);
void fhn_grl1(const double *, double *, double, double, const double *);
void tp06_fe(const double *, double *, double, double, const double *);
void tp06_grl1(const double *, double *, double, double, const double *);
void torord_fe(const double *, double *, double, double, const double *);
void torord_grl1(const double *, double *, double, double, const double *);
extern const int fhn_num_states, fhn_num_params, tp06_num_states, tp06_num_params, torord_num_states, torord_num_params;

#define MAX_NS 64
#define MAX_NP 160

static int model_info(int model, int scheme, step_fn *fn, int *ns, int *np) {
  switch (model) {
    case 0: *fn = scheme ? fhn_grl1 : fhn_fe; *ns = fhn_num_states; *np = fhn_num_params; return 0;
    case 1: *fn = scheme ? tp06_grl1 : tp06_fe; *ns = tp06_num_states; *np = tp06_num_params; return 0;
    case 2: *fn = scheme ? torord_grl1 : torord_fe; *ns = torord_num_states; *np = torord_num_params; return 0;
  }
  return -1;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* odesolver.py:67-79 over (ns, ld) SoA states; params (np,) shared or (np, ldp) per node */
int oracle_ode_step(int model, int scheme, int64_t n, int64_t ld, double *states, const double *params, int per_node,
                    int64_t ldp, double t, double dt) {
  step_fn fn;
  int ns, np;
  if (model_info(model, scheme, &fn, &ns, &np)) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double yi[MAX_NS], yo[MAX_NS], pr[MAX_NP];
    for (int k = 0; k < ns; ++k) yi[k] = states[(int64_t)k * ld + i];
    const double *p = params;
    if (per_node) {
      for (int k = 0; k < np; ++k) pr[k] = params[(int64_t)k * ldp + i];
      p = pr;
    }
    fn(yi, yo, t, dt, p);
    for (int k = 0; k < ns; ++k) states[(int64_t)k * ld + i] = yo[k];
  }
  return 0;
}

static void spmv(int64_t n, const int64_t *ip, const int32_t *ix, const double *a, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int64_t e = ip[i]; e < ip[i + 1]; ++e) s += a[e] * x[ix[e]];
    y[i] = s;
  }
}

static double dot(int64_t n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

typedef struct {
  int64_t n;
  const int64_t *indptr;
  const int32_t *indices;
  const double *A, *B, *dinv; /* A = C_m Mass + dt theta K ; B = C_m Mass - dt (1-theta) K ; 1/diag(A) */
  double theta, rtol, atol;
  int max_it;
  int n_stim;
  const double *stim_load; /* n_stim dense load vectors of length n */
  const double *stim_t0, *stim_t1, *stim_amp;
} oracle_pde;

/* base_model.py:208-245 with KSPCG+Jacobi: returns iterations (>=0) */
int oracle_pde_step(const oracle_pde *P, double t0, double t1, const double *v_prev, double *x, double *work /* 5n */,
                    double *rnorm_out) {
  const int64_t n = P->n;
  double *b = work, *r = work + n, *z = work + 2 * n, *p = work + 3 * n, *q = work + 4 * n;
  const double dt = t1 - t0, t = t0 + P->theta * dt;
  spmv(n, P->indptr, P->indices, P->B, v_prev, b);
  for (int k = 0; k < P->n_stim; ++k) {
    if (t >= P->stim_t0[k] && t <= P->stim_t1[k] && P->stim_amp[k] != 0.0) {
      const double f = dt * P->stim_amp[k];
      const double *s = P->stim_load + (int64_t)k * n;
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < n; ++i) b[i] += f * s[i];
    }
  }
  double bn = 0.0, rz = 0.0;
#pragma omp parallel for reduction(+ : bn, rz) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    x[i] = 0.0;
    r[i] = b[i];
    z[i] = P->dinv[i] * b[i];
    p[i] = z[i];
    bn += z[i] * z[i];
    rz += r[i] * z[i];
  }
  const double ttol = fmax(P->rtol * sqrt(bn), P->atol);
  double rnorm = sqrt(bn);
  int its = 0;
  while (rnorm > ttol && its < P->max_it) {
    spmv(n, P->indptr, P->indices, P->A, p, q);
    const double alpha = rz / dot(n, p, q);
    double zz = 0.0, rz_new = 0.0;
#pragma omp parallel for reduction(+ : zz, rz_new) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      z[i] = P->dinv[i] * r[i];
      zz += z[i] * z[i];
      rz_new += r[i] * z[i];
    }
    ++its;
    rnorm = sqrt(zz);
    const double beta = rz_new / rz;
    rz = rz_new;
    if (rnorm <= ttol) break;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
  if (rnorm_out) *rnorm_out = rnorm;
  return its;
}

/* monodomain_solver.py:53-116, nsteps consecutive steps; v (n) is pde.state == v_ on entry and exit.
 * Returns total CG iterations, or <0 on error. */
int64_t oracle_split_steps(const oracle_pde *P, int model, int scheme, int v_index, int64_t ld, double *states,
                           const double *params, double *v, double t0, double dt, int64_t nsteps, double theta_split) {
  const int64_t n = P->n;
  double *work = (double *)malloc(sizeof(double) * 7 * (size_t)(n > 0 ? n : 1));
  if (!work) return -1;
  double *v_prev = work + 5 * n, *x = work + 6 * n;
  double *vrow = states + (int64_t)v_index * ld;
  int64_t total = 0;
  double t = t0;
  for (int64_t s = 0; s < nsteps; ++s) {
    const double tn = t + dt;
    if (oracle_ode_step(model, scheme, n, ld, states, params, 0, 0, t, theta_split * dt)) {
      free(work);
      return -1;
    }
    memcpy(v_prev, vrow, sizeof(double) * n); /* to_dolfin, ode_to_pde, assign_previous */
    total += oracle_pde_step(P, t, tn, v_prev, x, work, NULL);
    memcpy(vrow, x, sizeof(double) * n); /* pde_to_ode, from_dolfin */
    if (fabs(theta_split - 1.0) > 1e-8) /* Strang corrective step, :98-113 */
      oracle_ode_step(model, scheme, n, ld, states, params, 0, 0, t + theta_split * dt, (1.0 - theta_split) * dt);
    memcpy(v, vrow, sizeof(double) * n);
    t = tn;
  }
  free(work);
  return total;
}
