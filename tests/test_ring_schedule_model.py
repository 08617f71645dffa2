"""Model check (CPU) of the request schedule of the shared-memory rings in pde_cg_stream_kernel, MODE 2
(fenicsx-beat_b200/csrc/pde_kernels.cu: ring_sweep, its steady-state fast path and ring_issue_share; host side:
pde_build_dictionary's choice of ring capacity and depth).

The kernel keeps, per offset cluster c = [lo_c, hi_c], a ring of `cap` elements of the gathered vector that follows the CTA's
sweep over its rows in tiles of T = 1024.  The requests for tile j are issued once all warps are done with tile j - (depth + 1),
so while they land, tiles j - depth .. j - 1 may still be read.  The arithmetic is restated here line by line and checked,
for random row ranges, clusters and vector lengths, against what the consumer needs:

  coverage     every element a row of tile t gathers (row + off, off in a cluster, inside the vector) has been requested by a
               tile <= t (the consumer waits for the full barriers of tiles 0 .. t in order);
  no clobber   a request of tile j never lands on the ring slot of an element that one of the tiles j - depth .. j - 1 still
               needs (different element, same slot modulo cap);
  exactly once no element is requested twice within a sweep (the byte count handed to the mbarrier would be wrong);
  geometry     every bulk copy is 16-byte aligned at both ends, inside the vector's padded length, and does not wrap in the ring.
"""
import numpy as np
import pytest

T = 1024      # kRingTile
C = 1024      # kRingChunk
MIN_DEPTH, MAX_DEPTH, STAGES = 1, 6, 7
TABLE_SMEM = 32 * 16 * 24 + 32 * 12
STAGE_BYTES = T + T // 32


def ceil_c(x):
    return (x + C - 1) & ~(C - 1)


def floor_c(x):
    return (x >> 10) << 10  # arithmetic shift: floor for negatives too, as in the kernel


def host_ring_parameters(clusters):
    """pde_build_dictionary: capacity (power of two) and tiles requested ahead for a given clustering, None if it does not fit."""
    span = max(hi - lo + 1 for lo, hi in clusters)
    lg = 9
    while (1 << lg) < span + (MIN_DEPTH + 1) * T + C:
        lg += 1

    def fits(l):
        return len(clusters) * (8 << l) + TABLE_SMEM + STAGES * (STAGE_BYTES + 16) + 256 <= 218 * 1024

    while ((1 << lg) - span - C) // T - 1 < 3 and fits(lg + 1):
        lg += 1
    if not fits(lg):
        return None
    depth = min(MAX_DEPTH, ((1 << lg) - span - C) // T - 1)
    return 1 << lg, max(depth, MIN_DEPTH)


def requests_of_tile(j, row_b, row_e, clusters, n_src):
    """[(cluster, lo, hi)] element ranges tile j asks for: the fast path of ring_sweep for whole tiles after the first,
    ring_issue_share otherwise."""
    n_even = (n_src + 1) & ~1
    full_tiles = (row_e - row_b) // T
    out = []
    if 1 <= j < full_tiles:
        for c, (lo_c, hi_c) in enumerate(clusters):
            chunk0 = ceil_c(min(row_b + T, row_e) + hi_c) - T
            lo = chunk0 + j * T
            if 0 <= lo < n_even:
                out.append((c, lo, lo + min(C, n_even - lo)))
        return out
    tile_hi, prev_hi = min(row_b + (j + 1) * T, row_e), min(row_b + j * T, row_e)
    for c, (lo_c, hi_c) in enumerate(clusters):
        end = ceil_c(min(tile_hi + hi_c, n_even))
        beg = floor_c(row_b + lo_c) if j == 0 else ceil_c(min(prev_hi + hi_c, n_even))
        nch = (end - beg) >> 10 if end > beg else 0
        for k in range(nch):
            lo, hi = max(beg + k * C, 0), min(beg + (k + 1) * C, n_even)
            if hi > lo:
                out.append((c, lo, hi))
    return out


def needs_of_tile(t, row_b, row_e, clusters, n_src):
    """per cluster the closed range of vector elements the rows of tile t may gather (clipped to the vector)."""
    lo_row, hi_row = row_b + t * T, min(row_b + (t + 1) * T, row_e) - 1
    out = []
    for lo_c, hi_c in clusters:
        a, b = max(lo_row + lo_c, 0), min(hi_row + hi_c, n_src - 1)
        out.append((a, b) if b >= a else None)
    return out


def check_sweep(row_b, row_e, clusters, n_src, cap, depth):
    ntiles = (row_e - row_b + T - 1) // T
    n_even = (n_src + 1) & ~1
    loaded = [np.zeros(n_even + 2 * C, dtype=np.int32) - 1 for _ in clusters]  # element -> tile that requested it
    reqs = [requests_of_tile(j, row_b, row_e, clusters, n_src) for j in range(ntiles)]
    for j, rq in enumerate(reqs):
        for c, lo, hi in rq:
            assert lo % 2 == 0 and hi % 2 == 0 and 0 <= lo < hi <= n_even, (j, c, lo, hi)           # 16-byte aligned, in range
            assert (lo & (cap - 1)) + (hi - lo) <= cap, ("wraps", j, c, lo, hi)
            assert (loaded[c][lo:hi] == -1).all(), ("requested twice", j, c, lo, hi)
            loaded[c][lo:hi] = j
    for t in range(ntiles):
        for c, rng in enumerate(needs_of_tile(t, row_b, row_e, clusters, n_src)):
            if rng is None:
                continue
            a, b = rng
            who = loaded[c][a:b + 1]
            assert (who >= 0).all() and (who <= t).all(), ("coverage", t, c, a, b, who.min(), who.max())
    # no clobber: what tile j requests must not share a ring slot with a DIFFERENT element still needed by tiles j-depth .. j-1
    for j, rq in enumerate(reqs):
        for c, lo, hi in rq:
            for t in range(max(0, j - depth), j):
                rng = needs_of_tile(t, row_b, row_e, clusters, n_src)[c]
                if rng is None:
                    continue
                a, b = rng
                # elements g in [lo, hi) and g' in [a, b] with g' != g and g' == g (mod cap): since hi - lo <= cap and
                # b - a < cap it is enough to look at g' = g - cap (requests move forward)
                assert not (lo - cap <= b and hi - 1 - cap >= a), ("clobber", j, t, c, lo, hi, a, b)


def box_clusters(nx, ny):
    """offset clusters of the Kuhn-split box numbered x-fastest: planes z-1, z, z+1 (what the dictionary finds on the slab)."""
    plane = nx * ny
    return [(-plane - nx - 1, -plane), (-nx - 1, nx + 1), (plane, plane + nx + 1)]


@pytest.mark.parametrize("nx,ny,nz", [(401, 141, 61), (801, 281, 121), (1251, 438, 188), (101, 281, 121), (41, 15, 7), (21, 15, 7)])
def test_slab_sweeps(nx, ny, nz):
    clusters = box_clusters(nx, ny)
    par = host_ring_parameters(clusters)
    assert par is not None
    cap, depth = par
    assert cap >= max(hi - lo + 1 for lo, hi in clusters) + (depth + 1) * T + C
    n = nx * ny * nz
    n_slices = (n + 31) // 32
    workers = 147
    per_cta = -(-n_slices // workers)
    per_cta = -(-per_cta // 32) * 32
    rng = np.random.default_rng(nx)
    ctas = sorted(set([0, 1, workers // 2, workers - 2, workers - 1] + list(rng.integers(0, workers, 3))))
    for b in ctas:
        s_begin, s_end = min(b * per_cta, n_slices), min((b + 1) * per_cta, n_slices)
        row_b, row_e = s_begin * 32, min(s_end * 32, n)
        if row_e > row_b:
            check_sweep(row_b, row_e, clusters, n, cap, depth)


def test_random_sweeps():
    rng = np.random.default_rng(7)
    for _ in range(60):
        ncl = int(rng.integers(1, 5))
        centres = np.sort(rng.integers(-200000, 200000, ncl))
        centres[rng.integers(0, ncl)] = 0  # the diagonal's cluster
        centres = np.unique(centres)
        clusters = []
        for cen in centres:
            w = int(rng.integers(0, 3000))
            clusters.append((int(cen) - w, int(cen) + int(rng.integers(0, 3000))))
        # clusters must be disjoint and ordered (the host merges offsets closer than its gap threshold)
        ok = all(clusters[i][1] < clusters[i + 1][0] for i in range(len(clusters) - 1))
        par = host_ring_parameters(clusters) if ok else None
        if par is None:
            continue
        cap, depth = par
        n = int(rng.integers(5000, 400000))
        row_b = int(rng.integers(0, max(1, n // T - 3))) * T
        row_e = min(n, row_b + int(rng.integers(1, 40 * T)))
        for d in sorted({1, depth}):
            check_sweep(row_b, row_e, clusters, n, cap, d)
