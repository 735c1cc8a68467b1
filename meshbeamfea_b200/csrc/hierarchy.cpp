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
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "mbfea_internal.hpp"

namespace mbfea {

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

static void build_level(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *pos, double cell,
                        HierLevel &L)
{
    L.n_fine = n;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t v = 0; v < n; ++v)
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], pos[3 * v + a]); hi[a] = std::max(hi[a], pos[3 * v + a]); }
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
    L.agg.assign((size_t)n, 0);
    L.agg_ptr.clear(); L.agg_rows.resize((size_t)n);
    int32_t na = 0;
    const double n_bricks = (double)cnt[0] * (double)cnt[1] * (double)cnt[2];
    if (n_bricks <= 8.0 * (double)n + 1024.0) {
        // counting sort by brick (stable: rows stay ascending inside a brick)
        const int64_t nbk = cnt[0] * cnt[1] * cnt[2];
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
    // coarse pattern: aggregate I couples to agg(col) of every block of its member rows.  Two passes (count,
    // fill) over ranges of aggregates, each thread with its own scratch.
    L.rowptr.assign((size_t)na + 1, 0);
    auto row_of = [&](int64_t I, std::vector<int32_t> &tmp) {
        tmp.clear();
        tmp.push_back((int32_t)I);
        for (int32_t q = L.agg_ptr[(size_t)I]; q < L.agg_ptr[(size_t)I + 1]; ++q) {
            const int32_t v = L.agg_rows[(size_t)q];
            for (int32_t k = rowptr[v]; k < rowptr[v + 1]; ++k) tmp.push_back(L.agg[(size_t)colidx[k]]);
        }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    };
    par_ranges(na, 1 << 10, [&](int64_t b0, int64_t b1) {
        std::vector<int32_t> tmp;
        for (int64_t I = b0; I < b1; ++I) { row_of(I, tmp); L.rowptr[(size_t)I + 1] = (int32_t)tmp.size(); }
    });
    for (int64_t I = 0; I < na; ++I) L.rowptr[(size_t)I + 1] += L.rowptr[(size_t)I];
    L.colidx.resize((size_t)L.rowptr[(size_t)na]);
    par_ranges(na, 1 << 10, [&](int64_t b0, int64_t b1) {
        std::vector<int32_t> tmp;
        for (int64_t I = b0; I < b1; ++I) { row_of(I, tmp); std::copy(tmp.begin(), tmp.end(), L.colidx.begin() + L.rowptr[(size_t)I]); }
    });
    // Galerkin contributor lists: fine block k = (a, b) is summed into coarse block (agg a, agg b) as
    // B(off a)^T A(a, b) B(off b); per coarse block the fine blocks are listed in ascending k.  The lists of coarse
    // row I come from I's member rows only (ascending rows => ascending k), so aggregates are independent: count,
    // prefix, fill -- both passes parallel over aggregates.
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
        build_level(n, rp.data(), ci.data(), pos.data(), cell, L);
        if (L.n_coarse >= n) break;   // no coarsening at this cell size (isolated points): stop
        pos = L.centroid; rp = L.rowptr; ci = L.colidx; n = L.n_coarse;
        H.levels.push_back(std::move(L));
        cell *= ratio;
    }
}

}  // namespace mbfea
