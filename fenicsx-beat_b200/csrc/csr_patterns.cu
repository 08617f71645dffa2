// Host-side set-up helper of the C ABI: classify the owned rows of the (mass, stiffness) CSR pair by their STENCIL, i.e.
// by (column offsets relative to the row, mass values, stiffness values) compared bit for bit.  On the structured meshes
// of the reference's benchmark (Niederer slab, `demos/niederer_benchmark.py:232-252`) every interior row carries the same
// 15 entries: 27 stencils cover a whole single-rank box and >98 % of a partitioned one, so the matrix stream of the CG
// iteration (180 of its 276 B/row) is redundant for those rows.  This routine finds the dictionary;
// no device code.
//
// Deterministic whatever the thread count: patterns are numbered by frequency (most rows first), ties by the first row
// they occur in.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mono_abi.h"
#include "host_threads.h"
#include "mono_ctx.h"

namespace {

struct Csr {
    const int64_t* indptr;
    const int32_t* indices;
    const double* mass;
    const double* stiff;
};

inline uint64_t mix(uint64_t h, uint64_t v) {  // splitmix-style avalanche of one more word
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xbf58476d1ce4e5b9ull;
    return h ^ (h >> 31);
}

uint64_t row_hash(const Csr& m, int64_t r) {
    const int64_t a = m.indptr[r], b = m.indptr[r + 1];
    uint64_t h = mix(0x243f6a8885a308d3ull, (uint64_t)(b - a));
    for (int64_t k = a; k < b; ++k) {
        uint64_t vm, vk;
        std::memcpy(&vm, m.mass + k, 8);
        std::memcpy(&vk, m.stiff + k, 8);
        h = mix(mix(mix(h, (uint64_t)(int64_t)(m.indices[k] - r)), vm), vk);
    }
    return h;
}

bool same_stencil(const Csr& m, int64_t r, int64_t s) {
    const int64_t a = m.indptr[r], b = m.indptr[s], w = m.indptr[r + 1] - a;
    if (m.indptr[s + 1] - b != w) return false;
    for (int64_t k = 0; k < w; ++k)
        if (m.indices[a + k] - r != m.indices[b + k] - s) return false;
    return std::memcmp(m.mass + a, m.mass + b, (size_t)w * 8) == 0 && std::memcmp(m.stiff + a, m.stiff + b, (size_t)w * 8) == 0;
}

struct Pattern {
    int64_t first_row;
    int64_t rows;
};

// hash -> indices of the patterns with that hash (collisions are resolved by comparing the rows themselves)
using Table = std::unordered_map<uint64_t, std::vector<int>>;

int find_or_add(const Csr& m, Table& tab, std::vector<Pattern>& pats, int64_t r, uint64_t h, int64_t count, size_t cap) {
    auto& bucket = tab[h];
    for (int p : bucket)
        if (same_stencil(m, pats[(size_t)p].first_row, r)) {
            pats[(size_t)p].rows += count;
            pats[(size_t)p].first_row = std::min(pats[(size_t)p].first_row, r);
            return p;
        }
    if (pats.size() >= cap) return -1;
    bucket.push_back((int)pats.size());
    pats.push_back({r, count});
    return (int)pats.size() - 1;
}

}  // namespace

extern "C" int mono_csr_row_patterns(int64_t n_owned, const int64_t* indptr, const int32_t* indices, const double* mass, const double* stiff,
                                     int max_patterns, uint8_t* pattern_of_row, int32_t* n_patterns, int64_t* representative_row,
                                     int64_t* rows_per_pattern) {
    if (n_owned < 0 || !indptr || (n_owned > 0 && (!indices || !mass || !stiff)) || max_patterns < 1 || max_patterns > 255 ||
        !pattern_of_row || !n_patterns || !representative_row || !rows_per_pattern)
        return mono_fail(nullptr, MONO_E_INVALID, "mono_csr_row_patterns: invalid argument");
    for (int64_t r = 0; r < n_owned; ++r)
        if (indptr[r + 1] < indptr[r]) return mono_fail(nullptr, MONO_E_INVALID, "mono_csr_row_patterns: indptr is not non-decreasing");
    try {
        const Csr m{indptr, indices, mass, stiff};
        const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), n_owned / 4096 + 1));
        constexpr size_t kLocalCap = 1 << 16;  // distinct stencils one thread tracks; rows beyond that count as general
        std::vector<std::vector<Pattern>> local((size_t)nt);
        std::vector<std::vector<int32_t>> local_id((size_t)nt);  // per row of the thread's range: local pattern or -1
        run_threads(nt, [&](int t) {
            const int64_t r0 = n_owned * t / nt, r1 = n_owned * (t + 1) / nt;
            Table tab;
            auto& ids = local_id[(size_t)t];
            ids.resize((size_t)(r1 - r0));
            for (int64_t r = r0; r < r1; ++r) ids[(size_t)(r - r0)] = find_or_add(m, tab, local[(size_t)t], r, row_hash(m, r), 1, kLocalCap);
        });
        // merge the per-thread dictionaries (thread order = row order, so first_row stays the global minimum)
        Table tab;
        std::vector<Pattern> pats;
        std::vector<std::vector<int>> to_global((size_t)nt);
        for (int t = 0; t < nt; ++t) {
            to_global[(size_t)t].resize(local[(size_t)t].size());
            for (size_t p = 0; p < local[(size_t)t].size(); ++p) {
                const Pattern& lp = local[(size_t)t][p];
                to_global[(size_t)t][p] = find_or_add(m, tab, pats, lp.first_row, row_hash(m, lp.first_row), lp.rows, (size_t)nt * kLocalCap);
            }
        }
        // keep the max_patterns most frequent ones
        std::vector<int> order(pats.size());
        for (size_t p = 0; p < pats.size(); ++p) order[p] = (int)p;
        std::sort(order.begin(), order.end(), [&](int a, int b) {
            return pats[(size_t)a].rows != pats[(size_t)b].rows ? pats[(size_t)a].rows > pats[(size_t)b].rows
                                                               : pats[(size_t)a].first_row < pats[(size_t)b].first_row;
        });
        const int keep = (int)std::min<size_t>(order.size(), (size_t)max_patterns);
        std::vector<int> final_id(pats.size(), 255);
        for (int k = 0; k < keep; ++k) {
            final_id[(size_t)order[(size_t)k]] = k;
            representative_row[k] = pats[(size_t)order[(size_t)k]].first_row;
            rows_per_pattern[k] = pats[(size_t)order[(size_t)k]].rows;
        }
        *n_patterns = keep;
        run_threads(nt, [&](int t) {
            const int64_t r0 = n_owned * t / nt, r1 = n_owned * (t + 1) / nt;
            for (int64_t r = r0; r < r1; ++r) {
                const int32_t lp = local_id[(size_t)t][(size_t)(r - r0)];
                pattern_of_row[r] = lp < 0 ? 255 : (uint8_t)final_id[(size_t)to_global[(size_t)t][(size_t)lp]];
            }
        });
    } catch (const std::exception& e) {
        return mono_fail(nullptr, MONO_E_NOMEM, std::string("mono_csr_row_patterns: ") + e.what());
    }
    return MONO_OK;
}
