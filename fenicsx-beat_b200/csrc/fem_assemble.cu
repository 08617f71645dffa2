// Host-side set-up helper of the C ABI: P1 simplex assembly of the OWNED rows of the mass and stiffness matrices
// (the two constant CSR matrices mono_pde_set_matrices takes).  No device code: this is the once-per-run work the
// reference leaves to dolfinx (`beat/monodomain_model.py:96-118`, forms v*w*dx and inner(M grad v, grad w)*dx); it lives
// here so that a 10^7-dof rank is assembled in seconds on the host cores instead of minutes in NumPy.
//
// Algorithm (race-free, deterministic, no global sort):
//   1. vertex -> incident (cell, local corner) lists for the owned vertices: a parallel counting sort over the cell
//      array (atomic counters), each short list sorted afterwards so the summation order does not depend on timing;
//   2. one row at a time (threads take chunks of rows): gather the row's <= ~30 distinct columns from its incident
//      cells, computing only row `a` of each element matrix, then sort the few columns.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mono_abi.h"
#include "host_threads.h"
#include "mono_ctx.h"

namespace {

constexpr int kMaxRow = 512;  // distinct columns of one row the gather can hold

// Row `a` of the element matrices of one P1 simplex: me[b] = |K| (1 + delta_ab) / ((D+1)(D+2)),
// ke[b] = |K| grad(phi_a) . M grad(phi_b).  Returns false for a degenerate cell.
template <int D>
bool element_row(const double* const xs[D + 1], int a, int m_kind, const double* M, double* me, double* ke) {
    double e[D][D];  // edge vectors from corner 0
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) e[i][j] = xs[i + 1][j] - xs[0][j];
    double g[D + 1][D];
    double det;
    if constexpr (D == 1) {
        det = e[0][0];
        g[1][0] = 1.0 / det;
    } else if constexpr (D == 2) {
        det = e[0][0] * e[1][1] - e[0][1] * e[1][0];
        g[1][0] = e[1][1] / det;
        g[1][1] = -e[1][0] / det;
        g[2][0] = -e[0][1] / det;
        g[2][1] = e[0][0] / det;
    } else {
        const double *p = e[0], *q = e[1], *r = e[2];
        double qr[3] = {q[1] * r[2] - q[2] * r[1], q[2] * r[0] - q[0] * r[2], q[0] * r[1] - q[1] * r[0]};
        double rp[3] = {r[1] * p[2] - r[2] * p[1], r[2] * p[0] - r[0] * p[2], r[0] * p[1] - r[1] * p[0]};
        double pq[3] = {p[1] * q[2] - p[2] * q[1], p[2] * q[0] - p[0] * q[2], p[0] * q[1] - p[1] * q[0]};
        det = p[0] * qr[0] + p[1] * qr[1] + p[2] * qr[2];
        for (int j = 0; j < 3; ++j) {
            g[1][j] = qr[j] / det;
            g[2][j] = rp[j] / det;
            g[3][j] = pq[j] / det;
        }
    }
    if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) return false;
    for (int j = 0; j < D; ++j) {
        double s = 0.0;
        for (int i = 1; i <= D; ++i) s += g[i][j];
        g[0][j] = -s;
    }
    constexpr double fact[4] = {1.0, 1.0, 2.0, 6.0};
    const double vol = std::fabs(det) / fact[D];
    double ga[D];  // grad(phi_a)^T M
    if (m_kind == 0) {
        for (int j = 0; j < D; ++j) ga[j] = M[0] * g[a][j];
    } else {
        for (int j = 0; j < D; ++j) {
            double s = 0.0;
            for (int i = 0; i < D; ++i) s += g[a][i] * M[i * D + j];
            ga[j] = s;
        }
    }
    const double mref = vol / double((D + 1) * (D + 2));
    for (int b = 0; b <= D; ++b) {
        double s = 0.0;
        for (int j = 0; j < D; ++j) s += ga[j] * g[b][j];
        ke[b] = vol * s;
        me[b] = (a == b) ? 2.0 * mref : mref;
    }
    return true;
}

template <int D>
int assemble(int64_t n_local, int64_t n_owned, int64_t n_cells, const int64_t* cells, const double* x, int x_ld, int m_kind,
             const double* M, int64_t* indptr, int32_t* indices, double* mass, double* stiff, std::string& err) {
    constexpr int NV = D + 1;
    const int nt = host_threads();
    const bool dbg = std::getenv("MONO_FEM_DEBUG") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        auto t1 = std::chrono::steady_clock::now();
        if (dbg) std::fprintf(stderr, "[fem_assemble nt=%d] %s %.3f s\n", nt, what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    // ---- 1. incident (cell, corner) lists of the owned vertices -----------------------------------------------
    // (parallel over CELL ranges with relaxed atomic counters; the order in which a row's incidences arrive is made
    // deterministic again by sorting each short list before it is used)
    std::vector<int64_t> vptr((size_t)n_owned + 1, 0);
    std::atomic<int> bad{0};
    const int64_t n_ent = n_cells * NV;
    run_threads(nt, [&](int t) {
        for (int64_t i = n_ent * t / nt, e = n_ent * (t + 1) / nt; i < e; ++i) {
            const int64_t v = cells[i];
            if (v < 0 || v >= n_local) {
                bad.store(1);
                return;
            }
            if (v < n_owned) __atomic_fetch_add(&vptr[(size_t)v + 1], (int64_t)1, __ATOMIC_RELAXED);
        }
    });
    if (bad.load()) {
        err = "mono_fem_assemble_p1: cell vertex index outside [0, n_local)";
        return MONO_E_INVALID;
    }
    lap("count incidences");
    for (int64_t v = 0; v < n_owned; ++v) vptr[(size_t)v + 1] += vptr[(size_t)v];
    std::vector<int64_t> inc((size_t)vptr[(size_t)n_owned]);
    {
        std::vector<int64_t> cur(vptr.begin(), vptr.end() - 1);
        run_threads(nt, [&](int t) {
            for (int64_t i = n_ent * t / nt, e = n_ent * (t + 1) / nt; i < e; ++i) {
                const int64_t v = cells[i];
                if (v < n_owned) inc[(size_t)__atomic_fetch_add(&cur[(size_t)v], (int64_t)1, __ATOMIC_RELAXED)] = i;  // cell * NV + corner
            }
        });
    }
    lap("fill incidences");
    // ---- 2. rows ---------------------------------------------------------------------------------------------
    const bool fill = indices != nullptr;
    std::atomic<int64_t> next{0};
    std::atomic<int> status{0};
    const int64_t chunk = 4096;
    run_threads(nt, [&](int) {
        int32_t col[kMaxRow];
        double vm[kMaxRow], vk[kMaxRow];
        int ord[kMaxRow];
        for (;;) {
            const int64_t b0 = next.fetch_add(chunk);
            if (b0 >= n_owned || status.load()) return;
            const int64_t b1 = std::min(n_owned, b0 + chunk);
            for (int64_t r = b0; r < b1; ++r) {
                int n = 0;
                std::sort(inc.begin() + vptr[(size_t)r], inc.begin() + vptr[(size_t)r + 1]);
                for (int64_t q = vptr[(size_t)r]; q < vptr[(size_t)r + 1]; ++q) {
                    const int64_t c = inc[(size_t)q] / NV;
                    const int a = (int)(inc[(size_t)q] % NV);
                    const int64_t* cv = cells + c * NV;
                    double me[NV], ke[NV];
                    if (fill) {
                        const double* xs[NV];
                        for (int b = 0; b < NV; ++b) xs[b] = x + cv[b] * x_ld;
                        const double* Mc = m_kind == 2 ? M + c * D * D : M;
                        if (!element_row<D>(xs, a, m_kind, Mc, me, ke)) {
                            status.store(1);
                            return;
                        }
                    }
                    for (int b = 0; b < NV; ++b) {
                        const int32_t cb = (int32_t)cv[b];
                        int s = 0;
                        while (s < n && col[s] != cb) ++s;
                        if (s == n) {
                            if (n == kMaxRow) {
                                status.store(2);
                                return;
                            }
                            col[n] = cb;
                            vm[n] = vk[n] = 0.0;
                            ++n;
                        }
                        if (fill) {
                            vm[s] += me[b];
                            vk[s] += ke[b];
                        }
                    }
                }
                if (!fill) {
                    indptr[r + 1] = n;
                    continue;
                }
                if (indptr[r + 1] - indptr[r] != n) {
                    status.store(3);
                    return;
                }
                for (int s = 0; s < n; ++s) ord[s] = s;
                std::sort(ord, ord + n, [&](int p, int q) { return col[p] < col[q]; });
                const int64_t o = indptr[r];
                for (int s = 0; s < n; ++s) {
                    indices[o + s] = col[ord[s]];
                    mass[o + s] = vm[ord[s]];
                    stiff[o + s] = vk[ord[s]];
                }
            }
        }
    });
    lap(fill ? "rows (values)" : "rows (pattern)");
    switch (status.load()) {
        case 1: err = "mono_fem_assemble_p1: degenerate cell (zero or non-finite volume)"; return MONO_E_INVALID;
        case 2: err = "mono_fem_assemble_p1: a row has more than 512 distinct columns"; return MONO_E_INVALID;
        case 3: err = "mono_fem_assemble_p1: indptr does not match the mesh (call with indices == NULL first)"; return MONO_E_INVALID;
        default: break;
    }
    if (!fill) {
        indptr[0] = 0;
        for (int64_t r = 0; r < n_owned; ++r) indptr[r + 1] += indptr[r];
    }
    return MONO_OK;
}

}  // namespace

extern "C" int mono_fem_assemble_p1(int tdim, int64_t n_local, int64_t n_owned, int64_t n_cells, const int64_t* cells, const double* x,
                                    int x_ld, int m_kind, const double* M, int64_t* indptr, int32_t* indices, double* mass,
                                    double* stiff) {
    std::string err;
    int rc = MONO_OK;
    if (tdim < 1 || tdim > 3 || n_owned < 0 || n_local < n_owned || n_cells < 0 || x_ld < tdim || m_kind < 0 || m_kind > 2 || !indptr ||
        (n_cells && (!cells || !x)) || !M || (indices && (!mass || !stiff)) || n_local > INT32_MAX) {
        err = "mono_fem_assemble_p1: invalid argument";
        rc = MONO_E_INVALID;
    } else {
        try {
            if (tdim == 1) rc = assemble<1>(n_local, n_owned, n_cells, cells, x, x_ld, m_kind, M, indptr, indices, mass, stiff, err);
            if (tdim == 2) rc = assemble<2>(n_local, n_owned, n_cells, cells, x, x_ld, m_kind, M, indptr, indices, mass, stiff, err);
            if (tdim == 3) rc = assemble<3>(n_local, n_owned, n_cells, cells, x, x_ld, m_kind, M, indptr, indices, mass, stiff, err);
        } catch (const std::exception& e) {
            err = std::string("mono_fem_assemble_p1: ") + e.what();
            rc = MONO_E_NOMEM;
        }
    }
    return rc == MONO_OK ? rc : mono_fail(nullptr, rc, err);
}
