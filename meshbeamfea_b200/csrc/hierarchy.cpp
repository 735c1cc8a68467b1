// Host half of the multilevel (rigid-body aggregate) preconditioner, SURVEY 8(f) rank 4: the level
// hierarchy of a one-rank job.  Nothing here is on the solver path yet -- the device half (Galerkin
// assembly from the contributor lists, restriction / prolongation, coarse solves) is the next round's
// work; this file fixes the data layout those kernels will consume and is pinned by CPU tests against
// scipy's P^T K P (tests/test_host_logic.py::test_aggregate_hierarchy_matches_galerkin_products).
//
// Level l+1 rows are AGGREGATES of level-l rows: bricks of a regular grid over the bounding box (cell
// edge `cell0` on level 0, times `ratio` per level).  An aggregate carries 6 degrees of freedom -- the
// rigid-body modes of its members about their centroid: a member at offset d = x - centroid moves by
// u = t + w x d, theta = w, i.e. its 6x6 prolongation block is B(d) = [[I, -[d]x], [0, I]].
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "mbfea_internal.hpp"

namespace mbfea {

namespace {
struct Lap {   // MBFEA_PREP_TRACE=1: wall time of the passes of the hierarchy build
    bool on = std::getenv("MBFEA_PREP_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void operator()(const char *what, int64_t n)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mbfea hierarchy] n=%-9lld %-28s %8.2f ms\n", (long long)n, what,
                     std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};
}  // namespace

template <class Fn>
static void par_ranges(int64_t n, int64_t min_per_thread, Fn fn)
{
    const unsigned hw = std::thread::hardware_concurrency();
    const int64_t nt = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw ? hw : 1, 32), n / std::max<int64_t>(1, min_per_thread)));
    if (nt <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int64_t t = 0; t < nt; ++t) th.emplace_back([=] { fn(n * t / nt, n * (t + 1) / nt); });
    for (auto &x : th) x.join();
}

// part_bounds (optional, nparts + 1 ascending row bounds): aggregates never straddle a part -- a brick cut by a
// partition boundary becomes one aggregate per part -- and are numbered part-major, so that every part owns a
// contiguous range of aggregates (multi-GPU: the fine-level transfer is then local to a rank).
// the same with the thread's index and the number of ranges actually used (per-thread scratch)
template <class Fn>
static int64_t par_ranges_t(int64_t n, int64_t min_per_thread, Fn fn)
{
    const unsigned hw = std::thread::hardware_concurrency();
    const int64_t nt = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw ? hw : 1, 32), n / std::max<int64_t>(1, min_per_thread)));
    if (nt <= 1) { fn((int64_t)0, n, (int64_t)0); return 1; }
    std::vector<std::thread> th;
    for (int64_t t = 0; t < nt; ++t) th.emplace_back([=] { fn(n * t / nt, n * (t + 1) / nt, t); });
    for (auto &x : th) x.join();
    return nt;
}
static int64_t par_ranges_count(int64_t n, int64_t min_per_thread)
{
    const unsigned hw = std::thread::hardware_concurrency();
    return std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw ? hw : 1, 32), n / std::max<int64_t>(1, min_per_thread)));
}

// First half of a level: the aggregates (needs the positions only, not the pattern).
static void build_level_aggregates(int64_t n, const double *pos, double cell, HierLevel &L, const int64_t *part_bounds = nullptr,
                                   int nparts = 1)
{
    Lap lap;
    L.n_fine = n;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    {   // bounding box: per-thread boxes, merged (min / max: the result does not depend on the split)
        const int64_t nt = par_ranges_count(n, 1 << 16);
        std::vector<double> box((size_t)nt * 6);
        par_ranges_t(n, 1 << 16, [&](int64_t b0, int64_t b1, int64_t t) {
            double l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int64_t v = b0; v < b1; ++v)
                for (int a = 0; a < 3; ++a) { l[a] = std::min(l[a], pos[3 * v + a]); h[a] = std::max(h[a], pos[3 * v + a]); }
            for (int a = 0; a < 3; ++a) { box[(size_t)t * 6 + a] = l[a]; box[(size_t)t * 6 + 3 + a] = h[a]; }
        });
        for (int64_t t = 0; t < nt; ++t)
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], box[(size_t)t * 6 + a]); hi[a] = std::max(hi[a], box[(size_t)t * 6 + 3 + a]); }
    }
    int64_t cnt[3];
    for (int a = 0; a < 3; ++a) cnt[a] = std::max<int64_t>(1, (int64_t)std::floor((hi[a] - lo[a]) / cell + 1e-9) + 1);
    // brick of every row; aggregates numbered in ascending (i, j, k) order of the occupied bricks, members in
    // ascending row order
    std::vector<int64_t> brick((size_t)n);
    par_ranges(n, 1 << 15, [&](int64_t b0, int64_t b1) {
        for (int64_t v = b0; v < b1; ++v) {
            int64_t c[3];
            for (int a = 0; a < 3; ++a)
                c[a] = std::min<int64_t>(cnt[a] - 1, std::max<int64_t>(0, (int64_t)std::floor((pos[3 * v + a] - lo[a]) / cell + 1e-9)));
            brick[(size_t)v] = (c[0] * cnt[1] + c[1]) * cnt[2] + c[2];
        }
    });
    int64_t n_keys = cnt[0] * cnt[1] * cnt[2];
    if (part_bounds != nullptr && nparts > 1) {
        const int64_t per_part = n_keys;
        for (int q = 0; q < nparts; ++q)
            for (int64_t v = part_bounds[q]; v < part_bounds[q + 1]; ++v) brick[(size_t)v] += (int64_t)q * per_part;
        n_keys *= nparts;
    }
    lap("bricks", n);
    L.agg.assign((size_t)n, 0);
    L.agg_ptr.clear(); L.agg_rows.resize((size_t)n);
    int32_t na = 0;
    const double n_bricks = (double)cnt[0] * (double)cnt[1] * (double)cnt[2] * (double)((part_bounds && nparts > 1) ? nparts : 1);
    if (n_bricks <= 8.0 * (double)n + 1024.0) {
        // counting sort by brick (stable: rows stay ascending inside a brick)
        const int64_t nbk = n_keys;
        std::vector<int32_t> start((size_t)nbk + 1, 0);
        for (int64_t v = 0; v < n; ++v) start[(size_t)brick[(size_t)v] + 1]++;
        std::vector<int32_t> id((size_t)nbk, -1);
        for (int64_t b = 0; b < nbk; ++b) {
            if (start[(size_t)b + 1] > 0) { id[(size_t)b] = na++; L.agg_ptr.push_back(start[(size_t)b]); }
            start[(size_t)b + 1] += start[(size_t)b];
        }
        // start[b] is now the first slot of brick b (prefix of counts): agg_ptr collected the raw counts' prefix
        L.agg_ptr.clear();
        for (int64_t b = 0; b < nbk; ++b)
            if (id[(size_t)b] >= 0) L.agg_ptr.push_back(start[(size_t)b]);
        std::vector<int32_t> cur(start.begin(), start.end() - 1);
        for (int64_t v = 0; v < n; ++v) {
            const int64_t b = brick[(size_t)v];
            L.agg[(size_t)v] = id[(size_t)b];
            L.agg_rows[(size_t)cur[(size_t)b]++] = (int32_t)v;
        }
    } else {
        std::vector<std::pair<int64_t, int32_t>> key((size_t)n);
        for (int64_t v = 0; v < n; ++v) key[(size_t)v] = {brick[(size_t)v], (int32_t)v};
        std::sort(key.begin(), key.end());
        for (int64_t i = 0; i < n; ++i) {
            if (i == 0 || key[(size_t)i].first != key[(size_t)i - 1].first) { L.agg_ptr.push_back((int32_t)i); ++na; }
            L.agg[(size_t)key[(size_t)i].second] = na - 1;
            L.agg_rows[(size_t)i] = key[(size_t)i].second;
        }
    }
    L.agg_ptr.push_back((int32_t)n);
    L.n_coarse = na;
    lap("aggregate lists", n);
    L.centroid.assign((size_t)3 * na, 0.0);
    par_ranges(na, 1 << 12, [&](int64_t b0, int64_t b1) {
        for (int64_t I = b0; I < b1; ++I) {
            double s[3] = {0, 0, 0};
            for (int32_t q = L.agg_ptr[(size_t)I]; q < L.agg_ptr[(size_t)I + 1]; ++q)   // fixed order: ascending rows
                for (int a = 0; a < 3; ++a) s[a] += pos[3 * (int64_t)L.agg_rows[(size_t)q] + a];
            const double m = (double)(L.agg_ptr[(size_t)I + 1] - L.agg_ptr[(size_t)I]);
            for (int a = 0; a < 3; ++a) L.centroid[(size_t)3 * I + a] = s[a] / m;
        }
    });
    L.off.resize((size_t)3 * n);
    par_ranges(n, 1 << 15, [&](int64_t b0, int64_t b1) {
        for (int64_t v = b0; v < b1; ++v)
            for (int a = 0; a < 3; ++a) L.off[(size_t)3 * v + a] = pos[3 * v + a] - L.centroid[(size_t)3 * L.agg[(size_t)v] + a];
    });
    lap("centroids + offsets", n);
}

// Second half (fine level only; coarse-to-coarse transfers build their own patterns in build_transfer): the coarse
// pattern and the Galerkin contributor lists from the fine pattern.
static void build_level_pattern(int64_t n, const int32_t *rowptr, const int32_t *colidx, HierLevel &L)
{
    Lap lap;
    const int64_t na = L.n_coarse;
    // coarse pattern: aggregate I couples to agg(col) of every block of its member rows.  Every thread generates the
    // rows of its contiguous range of aggregates once, into one flat buffer; after the prefix sum the buffer is exactly
    // the slice [rowptr[first], rowptr[last + 1]) of the column array.
    L.rowptr.assign((size_t)na + 1, 0);
    {
        const int64_t nt = par_ranges_count(na, 1 << 10);
        std::vector<std::vector<int32_t>> flat((size_t)nt);
        par_ranges_t(na, 1 << 10, [&](int64_t b0, int64_t b1, int64_t t) {
            std::vector<int32_t> tmp;
            std::vector<int32_t> &mine = flat[(size_t)t];
            mine.reserve((size_t)(b1 - b0) * 28);
            for (int64_t I = b0; I < b1; ++I) {
                tmp.clear();
                tmp.push_back((int32_t)I);
                for (int32_t q = L.agg_ptr[(size_t)I]; q < L.agg_ptr[(size_t)I + 1]; ++q) {
                    const int32_t v = L.agg_rows[(size_t)q];
                    for (int32_t k = rowptr[v]; k < rowptr[v + 1]; ++k) tmp.push_back(L.agg[(size_t)colidx[k]]);
                }
                std::sort(tmp.begin(), tmp.end());
                tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                L.rowptr[(size_t)I + 1] = (int32_t)tmp.size();
                mine.insert(mine.end(), tmp.begin(), tmp.end());
            }
        });
        for (int64_t I = 0; I < na; ++I) L.rowptr[(size_t)I + 1] += L.rowptr[(size_t)I];
        L.colidx.resize((size_t)L.rowptr[(size_t)na]);
        par_ranges_t(na, 1 << 10, [&](int64_t b0, int64_t, int64_t t) {
            if (!flat[(size_t)t].empty())
                std::memcpy(L.colidx.data() + L.rowptr[(size_t)b0], flat[(size_t)t].data(), flat[(size_t)t].size() * sizeof(int32_t));
        });
    }
    lap("coarse pattern", n);
}

// Galerkin contributor lists of the fine level: fine block k = (a, b) is summed into coarse block (agg a, agg b) as
// B(off a)^T A(a, b) B(off b); per coarse block the fine blocks are listed in ascending k.  The lists of coarse
// row I come from I's member rows only (ascending rows => ascending k), so aggregates are independent: count,
// prefix, fill -- both passes parallel over aggregates.  Needs the coarse pattern; nothing below level 1 needs the
// lists, so build_ml_hierarchy runs this next to the coarser levels.
static void build_level_galerkin_lists(int64_t n, const int32_t *rowptr, const int32_t *colidx, HierLevel &L)
{
    Lap lap;
    const int64_t na = L.n_coarse;
    const int64_t nnz_f = rowptr[n], nnz_c = (int64_t)L.colidx.size();
    L.gal_ptr.assign((size_t)nnz_c + 1, 0);
    auto for_blocks_of = [&](int64_t I, auto &&visit) {   // visit(coarse block index, fine block index), fine ascending
        const int32_t *cb = L.colidx.data() + L.rowptr[(size_t)I], *ce = L.colidx.data() + L.rowptr[(size_t)I + 1];
        for (int32_t q = L.agg_ptr[(size_t)I]; q < L.agg_ptr[(size_t)I + 1]; ++q) {
            const int32_t v = L.agg_rows[(size_t)q];
            for (int32_t k = rowptr[v]; k < rowptr[v + 1]; ++k)
                visit((int32_t)(std::lower_bound(cb, ce, L.agg[(size_t)colidx[k]]) - L.colidx.data()), k);
        }
    };
    par_ranges(na, 1 << 10, [&](int64_t b0, int64_t b1) {
        for (int64_t I = b0; I < b1; ++I) for_blocks_of(I, [&](int32_t kc, int32_t) { L.gal_ptr[(size_t)kc + 1]++; });
    });
    for (int64_t k = 0; k < nnz_c; ++k) L.gal_ptr[(size_t)k + 1] += L.gal_ptr[(size_t)k];
    L.gal_src.resize((size_t)nnz_f);
    par_ranges(na, 1 << 10, [&](int64_t b0, int64_t b1) {
        std::vector<int32_t> cur;
        for (int64_t I = b0; I < b1; ++I) {
            const int32_t c0 = L.rowptr[(size_t)I], c1 = L.rowptr[(size_t)I + 1];
            cur.assign(L.gal_ptr.begin() + c0, L.gal_ptr.begin() + c1);
            for_blocks_of(I, [&](int32_t kc, int32_t k) { L.gal_src[(size_t)cur[(size_t)(kc - c0)]++] = k; });
        }
    });
    lap("galerkin lists", n);
}

// Symbolic structures of one coarse-to-coarse transfer (see MlTransfer).  (rowptr, colidx): pattern of level l's
// matrix (diagonal present), L: its aggregates.
static void build_transfer(int64_t n, const int32_t *rowptr, const int32_t *colidx, const HierLevel &L, bool smoothed,
                           MlTransfer &T)
{
    Lap lap;
    const int64_t nc = L.n_coarse;
    T.smoothed = smoothed; T.n_fine = n; T.n_coarse = nc;
    // rows of a pattern built in two parallel passes (count, fill) from a per-row generator that pushes candidate
    // columns through `add` (a per-thread marker array removes duplicates before the row is sorted)
    struct Dedup {
        std::vector<int32_t> mark, out;
        int32_t stamp = 0;
        void begin(int32_t row) { stamp = row; out.clear(); }
        void add(int32_t c) { if (mark[(size_t)c] != stamp) { mark[(size_t)c] = stamp; out.push_back(c); } }
    };
    auto build_rows = [&](int64_t rows, int64_t ncols, int64_t min_per_thread, hvec<int32_t> &rp, hvec<int32_t> &ci, auto &&gen) {
        rp.assign((size_t)rows + 1, 0);
        // every thread generates the rows of its contiguous range once, into one flat buffer; after the prefix sum the
        // buffer is exactly the slice [rp[first row], rp[last row + 1]) of the column array
        const int64_t nt = par_ranges_count(rows, min_per_thread);
        std::vector<std::vector<int32_t>> flat((size_t)nt);
        par_ranges_t(rows, min_per_thread, [&](int64_t b0, int64_t b1, int64_t t) {
            Dedup d;
            d.mark.assign((size_t)ncols, -1);
            std::vector<int32_t> &mine = flat[(size_t)t];
            for (int64_t i = b0; i < b1; ++i) {
                d.begin((int32_t)i);
                gen(i, d);
                std::sort(d.out.begin(), d.out.end());
                rp[(size_t)i + 1] = (int32_t)d.out.size();
                mine.insert(mine.end(), d.out.begin(), d.out.end());
            }
        });
        for (int64_t i = 0; i < rows; ++i) rp[(size_t)i + 1] += rp[(size_t)i];
        ci.resize((size_t)rp[(size_t)rows]);
        par_ranges_t(rows, min_per_thread, [&](int64_t b0, int64_t, int64_t t) {
            if (!flat[(size_t)t].empty()) std::memcpy(ci.data() + rp[(size_t)b0], flat[(size_t)t].data(), flat[(size_t)t].size() * sizeof(int32_t));
        });
    };
    // P: row i couples to agg(j) of every block (i, j) when smoothed, to agg(i) alone otherwise
    build_rows(n, nc, 1 << 10, T.prow, T.pcol, [&](int64_t i, Dedup &d) {
        d.add(L.agg[(size_t)i]);
        if (smoothed)
            for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) d.add(L.agg[(size_t)colidx[k]]);
    });
    T.pown.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i)
        T.pown[(size_t)i] = (int32_t)(std::lower_bound(T.pcol.begin() + T.prow[(size_t)i], T.pcol.begin() + T.prow[(size_t)i + 1],
                                                       L.agg[(size_t)i]) - T.pcol.begin());
    lap("P pattern", n);
    // transpose structure (counting sort by column keeps block indices, hence fine rows, ascending)
    const int64_t nnzp = (int64_t)T.pcol.size();
    T.rrow.assign((size_t)nc + 1, 0);
    for (int64_t b = 0; b < nnzp; ++b) T.rrow[(size_t)T.pcol[(size_t)b] + 1]++;
    for (int64_t I = 0; I < nc; ++I) T.rrow[(size_t)I + 1] += T.rrow[(size_t)I];
    T.rblk.resize((size_t)nnzp); T.rfine.resize((size_t)nnzp);
    {
        std::vector<int32_t> cur(T.rrow.begin(), T.rrow.end() - 1);
        for (int64_t i = 0; i < n; ++i)
            for (int32_t b = T.prow[(size_t)i]; b < T.prow[(size_t)i + 1]; ++b) {
                const int32_t slot = cur[(size_t)T.pcol[(size_t)b]]++;
                T.rblk[(size_t)slot] = b; T.rfine[(size_t)slot] = (int32_t)i;
            }
    }
    lap("P transpose", n);
    // A P: row i couples to the P columns of every k with a block (i, k)
    build_rows(n, nc, 64, T.aprow, T.apcol, [&](int64_t i, Dedup &d) {
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
            const int32_t j = colidx[k];
            for (int32_t b = T.prow[(size_t)j]; b < T.prow[(size_t)j + 1]; ++b) d.add(T.pcol[(size_t)b]);
        }
    });
    lap("A P pattern", n);
    // P^T (A P): coarse row I couples to the A P columns of every fine row with a P block in column I
    build_rows(nc, nc, 8, T.crow, T.ccol, [&](int64_t I, Dedup &d) {
        d.add((int32_t)I);
        for (int32_t e = T.rrow[(size_t)I]; e < T.rrow[(size_t)I + 1]; ++e) {
            const int32_t i = T.rfine[(size_t)e];
            for (int32_t b = T.aprow[(size_t)i]; b < T.aprow[(size_t)i + 1]; ++b) d.add(T.apcol[(size_t)b]);
        }
    });
    lap("P^T A P pattern", n);
    T.cdiag.resize((size_t)nc);
    for (int64_t I = 0; I < nc; ++I)
        T.cdiag[(size_t)I] = (int32_t)(std::lower_bound(T.ccol.begin() + T.crow[(size_t)I], T.ccol.begin() + T.crow[(size_t)I + 1],
                                                        (int32_t)I) - T.ccol.begin());
}

void build_ml_hierarchy(int64_t nb, const int32_t *rowptr, const int32_t *colidx, const double *pos_rm, double cell0,
                        double ratio, int max_levels, int64_t coarsest_rows, int smooth_from, MlHierarchy &H,
                        const int64_t *part_bounds, int nparts, const FinePatternGate *fine_pattern)
{
    // Built IN PLACE: a hierarchy object that is handed in again (the next set_job of the same context) keeps the
    // capacity of every array, so a re-submitted mesh of similar size touches no fresh pages (first-touch page faults of
    // several hundred MB were a third of the build).
    size_t used = 0;
    std::thread galerkin_job;
    struct Join { std::thread &t; ~Join() { if (t.joinable()) t.join(); } } join_galerkin{galerkin_job};
    if (H.agg.size() < (size_t)std::max(1, max_levels)) { H.agg.resize((size_t)max_levels); H.xfer.resize((size_t)max_levels); }
    if (nb > 0 && cell0 > 0.0 && ratio > 1.0) {
        const double *pos = pos_rm;
        const int32_t *crp = rowptr, *cci = colidx;
        double cell = cell0;
        int64_t n = nb;
        for (int l = 0; l + 1 < max_levels && n > coarsest_rows; ++l) {
            HierLevel &L = H.agg[used];
            MlTransfer &T = H.xfer[used];
            build_level_aggregates(n, pos, cell, L, l == 0 ? part_bounds : nullptr, l == 0 ? nparts : 1);
            if (l == 0) {
                // the aggregates needed the positions only; the fine pattern may still be under construction on the
                // caller's side (set_job builds it on another thread): wait for it here
                if (fine_pattern && !(*fine_pattern)(crp, cci)) { used = 0; break; }
                build_level_pattern(n, crp, cci, L);
                // the contributor lists of the fine level are needed by nobody below: built next to the coarser levels
                // (H.agg was sized for every level up front, so `L` stays where it is)
                galerkin_job = std::thread([n, crp, cci, &L] { build_level_galerkin_lists(n, crp, cci, L); });
            } else {
                L.rowptr.clear(); L.colidx.clear(); L.gal_ptr.clear(); L.gal_src.clear();
            }
            if (L.n_coarse >= n) break;   // no coarsening at this cell size: stop
            if (l >= 1) {
                build_transfer(n, crp, cci, L, l >= smooth_from, T);
                crp = T.crow.data(); cci = T.ccol.data();
            } else {
                T = MlTransfer();
                crp = L.rowptr.data(); cci = L.colidx.data();
            }
            pos = L.centroid.data(); n = L.n_coarse;
            ++used;
            cell *= ratio;
        }
    }
    if (galerkin_job.joinable()) galerkin_job.join();
    H.agg.resize(used); H.xfer.resize(used);
}

// ---- flat image of a hierarchy (several ranks: rank 0 builds it with every host core, the others receive it) ----
bool ml_image_layout(const MlHierarchy &H, int world, int64_t nbg, const int32_t *fine_rowptr, MlImageLayout &lay)
{
    const size_t L = H.levels();
    if (L == 0 || L > (size_t)kMlImageMaxLevels || world < 1 || world > kMlImageMaxWorld) return false;
    MlImageHeader &hd = lay.hd;
    std::memset(&hd, 0, sizeof hd);
    hd.magic = kMlImageMagic; hd.levels = (int64_t)L; hd.world = world; hd.nbg = nbg; hd.fine_nnz = fine_rowptr[nbg];
    const HierLevel &h0 = H.agg[0];
    int32_t biggest = 0;
    for (int64_t I = 0; I < h0.n_coarse; ++I) biggest = std::max(biggest, h0.agg_ptr[(size_t)I + 1] - h0.agg_ptr[(size_t)I]);
    hd.group0 = biggest <= 8 ? 8 : 32;
    for (int q = 0; q < world; ++q) {   // aggregates are numbered rank-major
        const int64_t lo = nbg * q / world, hi = nbg * (q + 1) / world;
        int32_t top = -1;
        for (int64_t v = lo; v < hi; ++v) top = std::max(top, h0.agg[(size_t)v]);
        hd.agg_begin[q + 1] = std::max<int64_t>(hd.agg_begin[q], (int64_t)top + 1);
    }
    for (int q = 0; q <= world; ++q) {
        hd.own_first[q] = fine_rowptr[nbg * q / world];
        hd.blk_begin[q] = h0.rowptr[(size_t)hd.agg_begin[q]];
        hd.aggptr_at[q] = h0.agg_ptr[(size_t)hd.agg_begin[q]];
        hd.galptr_at[q] = h0.gal_ptr[(size_t)hd.blk_begin[q]];
    }
    // diagonal positions of the level-1 pattern
    hvec<int32_t> &diag1 = lay.diag1;
    diag1.resize((size_t)h0.n_coarse);
    for (int64_t I = 0; I < h0.n_coarse; ++I)
        diag1[(size_t)I] = (int32_t)(std::lower_bound(h0.colidx.begin() + h0.rowptr[(size_t)I], h0.colidx.begin() + h0.rowptr[(size_t)I + 1],
                                                      (int32_t)I) - h0.colidx.begin());
    struct Src { const void *p; size_t bytes; int64_t cnt; };
    std::vector<std::vector<Src>> src(L);
    auto vi = [](const auto &v) { return Src{v.data(), v.size() * sizeof(int32_t), (int64_t)v.size()}; };
    auto vd = [](const auto &v) { return Src{v.data(), v.size() * sizeof(double), (int64_t)v.size()}; };
    src[0] = {vi(h0.agg), vd(h0.off), vi(h0.agg_ptr), vi(h0.agg_rows), vi(h0.rowptr), vi(h0.colidx), vi(diag1), vi(h0.gal_ptr), vi(h0.gal_src)};
    for (size_t l = 1; l < L; ++l) {
        const HierLevel &a = H.agg[l];
        const MlTransfer &t = H.xfer[l];
        src[l] = {vi(a.agg), vd(a.off), vi(t.prow), vi(t.pcol), vi(t.rrow), vi(t.rblk), vi(t.rfine), vi(t.aprow), vi(t.apcol),
                  vi(t.crow), vi(t.ccol), vi(t.cdiag)};
    }
    size_t at = (sizeof(MlImageHeader) + 15) & ~(size_t)15;
    lay.parts.clear();
    for (size_t l = 0; l < L; ++l) {
        hd.lv[l].n_fine = H.agg[l].n_fine; hd.lv[l].n_coarse = H.agg[l].n_coarse;
        hd.lv[l].smoothed = (l >= 1 && H.xfer[l].smoothed) ? 1 : 0;
        for (size_t k = 0; k < src[l].size(); ++k) {
            hd.lv[l].off[k] = (int64_t)at; hd.lv[l].cnt[k] = src[l][k].cnt;
            if (src[l][k].bytes) lay.parts.push_back(MlImageLayout::Part{src[l][k].p, src[l][k].bytes, at});
            at += (src[l][k].bytes + 15) & ~(size_t)15;
        }
    }
    hd.total_bytes = (int64_t)at;
    return true;
}

bool ml_image_pack(const MlHierarchy &H, int world, int64_t nbg, const int32_t *fine_rowptr, std::vector<unsigned char> &out)
{
    MlImageLayout lay;
    if (!ml_image_layout(H, world, nbg, fine_rowptr, lay)) return false;
    out.resize((size_t)lay.hd.total_bytes);   // one allocation, every array copied once (in parallel)
    std::memcpy(out.data(), &lay.hd, sizeof lay.hd);
    std::vector<std::thread> th;
    for (const MlImageLayout::Part &pt : lay.parts) {
        unsigned char *dst = out.data() + pt.at;
        if (pt.bytes > (size_t)(8 << 20)) th.emplace_back([=] { std::memcpy(dst, pt.p, pt.bytes); });
        else std::memcpy(dst, pt.p, pt.bytes);
    }
    for (auto &x : th) x.join();
    return true;
}

void build_hierarchy(int64_t nb, const int32_t *rowptr, const int32_t *colidx, const double *pos_rm, double cell0,
                     double ratio, int max_levels, int64_t coarsest_rows, Hierarchy &H)
{
    H.levels.clear();
    if (nb <= 0 || !(cell0 > 0.0) || !(ratio > 1.0)) return;
    std::vector<double> pos(pos_rm, pos_rm + 3 * nb);
    std::vector<int32_t> rp(rowptr, rowptr + nb + 1), ci(colidx, colidx + rowptr[nb]);
    double cell = cell0;
    int64_t n = nb;
    for (int l = 1; l < max_levels && n > coarsest_rows; ++l) {
        HierLevel L;
        build_level_aggregates(n, pos.data(), cell, L);
        build_level_pattern(n, rp.data(), ci.data(), L);
        build_level_galerkin_lists(n, rp.data(), ci.data(), L);
        if (L.n_coarse >= n) break;   // no coarsening at this cell size (isolated points): stop
        pos = L.centroid; rp.assign(L.rowptr.begin(), L.rowptr.end()); ci.assign(L.colidx.begin(), L.colidx.end()); n = L.n_coarse;
        H.levels.push_back(std::move(L));
        cell *= ratio;
    }
}

}  // namespace mbfea
