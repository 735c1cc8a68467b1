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

#include "mbfea_internal.hpp"

namespace mbfea {

static void build_level(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *pos, double cell,
                        HierLevel &L)
{
    L.n_fine = n;
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t v = 0; v < n; ++v)
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], pos[3 * v + a]); hi[a] = std::max(hi[a], pos[3 * v + a]); }
    int64_t cnt[3];
    for (int a = 0; a < 3; ++a) cnt[a] = std::max<int64_t>(1, (int64_t)std::floor((hi[a] - lo[a]) / cell + 1e-9) + 1);
    // brick of every row; aggregates numbered in ascending (i, j, k) order of the occupied bricks
    std::vector<std::pair<int64_t, int32_t>> key((size_t)n);
    for (int64_t v = 0; v < n; ++v) {
        int64_t c[3];
        for (int a = 0; a < 3; ++a)
            c[a] = std::min<int64_t>(cnt[a] - 1, std::max<int64_t>(0, (int64_t)std::floor((pos[3 * v + a] - lo[a]) / cell + 1e-9)));
        key[(size_t)v] = {(c[0] * cnt[1] + c[1]) * cnt[2] + c[2], (int32_t)v};
    }
    std::sort(key.begin(), key.end());
    L.agg.assign((size_t)n, 0);
    L.agg_ptr.clear(); L.agg_rows.resize((size_t)n);
    int32_t na = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (i == 0 || key[(size_t)i].first != key[(size_t)i - 1].first) { L.agg_ptr.push_back((int32_t)i); ++na; }
        L.agg[(size_t)key[(size_t)i].second] = na - 1;
        L.agg_rows[(size_t)i] = key[(size_t)i].second;   // ascending row id inside an aggregate (pair sort)
    }
    L.agg_ptr.push_back((int32_t)n);
    L.n_coarse = na;
    L.centroid.assign((size_t)3 * na, 0.0);
    for (int32_t I = 0; I < na; ++I) {
        double s[3] = {0, 0, 0};
        for (int32_t q = L.agg_ptr[(size_t)I]; q < L.agg_ptr[(size_t)I + 1]; ++q)   // fixed order: ascending rows
            for (int a = 0; a < 3; ++a) s[a] += pos[3 * (int64_t)L.agg_rows[(size_t)q] + a];
        const double m = (double)(L.agg_ptr[(size_t)I + 1] - L.agg_ptr[(size_t)I]);
        for (int a = 0; a < 3; ++a) L.centroid[(size_t)3 * I + a] = s[a] / m;
    }
    L.off.resize((size_t)3 * n);
    for (int64_t v = 0; v < n; ++v)
        for (int a = 0; a < 3; ++a) L.off[(size_t)3 * v + a] = pos[3 * v + a] - L.centroid[(size_t)3 * L.agg[(size_t)v] + a];
    // coarse pattern: aggregate I couples to agg(col) of every block of its member rows
    L.rowptr.assign((size_t)na + 1, 0);
    L.colidx.clear();
    std::vector<int32_t> tmp;
    for (int32_t I = 0; I < na; ++I) {
        tmp.clear();
        tmp.push_back(I);
        for (int32_t q = L.agg_ptr[(size_t)I]; q < L.agg_ptr[(size_t)I + 1]; ++q) {
            const int32_t v = L.agg_rows[(size_t)q];
            for (int32_t k = rowptr[v]; k < rowptr[v + 1]; ++k) tmp.push_back(L.agg[(size_t)colidx[k]]);
        }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        L.colidx.insert(L.colidx.end(), tmp.begin(), tmp.end());
        L.rowptr[(size_t)I + 1] = (int32_t)L.colidx.size();
    }
    // Galerkin contributor lists: fine block k = (a, b) is summed into coarse block (agg a, agg b) as
    // B(off a)^T A(a, b) B(off b); per coarse block the fine blocks are listed in ascending k
    const int64_t nnz_f = rowptr[n], nnz_c = (int64_t)L.colidx.size();
    std::vector<int32_t> target((size_t)nnz_f);
    L.gal_ptr.assign((size_t)nnz_c + 1, 0);
    for (int64_t a = 0; a < n; ++a) {
        const int32_t I = L.agg[(size_t)a];
        const int32_t *cb = L.colidx.data() + L.rowptr[(size_t)I], *ce = L.colidx.data() + L.rowptr[(size_t)I + 1];
        for (int32_t k = rowptr[a]; k < rowptr[a + 1]; ++k) {
            const int32_t J = L.agg[(size_t)colidx[k]];
            const int32_t kc = (int32_t)(std::lower_bound(cb, ce, J) - L.colidx.data());
            target[(size_t)k] = kc;
            L.gal_ptr[(size_t)kc + 1]++;
        }
    }
    for (int64_t k = 0; k < nnz_c; ++k) L.gal_ptr[(size_t)k + 1] += L.gal_ptr[(size_t)k];
    L.gal_src.resize((size_t)nnz_f);
    std::vector<int32_t> cur(L.gal_ptr.begin(), L.gal_ptr.end() - 1);
    for (int64_t k = 0; k < nnz_f; ++k) L.gal_src[(size_t)cur[(size_t)target[(size_t)k]]++] = (int32_t)k;
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
