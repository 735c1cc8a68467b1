// Host driver of the multilevel preconditioner (see mlprec.hpp).
#include "mlprec.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"

namespace mbfea {

using namespace mbfea::dev;

namespace {
template <class T, class A>
void up(cudaStream_t st, DevBuf<T> &d, const std::vector<T, A> &h)
{
    CK(d.alloc(h.size()));
    if (!h.empty()) CK(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}
double wall_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
}  // namespace

void MlPrec::reset()
{
    symbolic_ok = numeric_ok = false;
    // the level objects (and their device buffers) are kept for the next job: a re-submitted mesh of the same size then
    // costs no cudaFree / cudaMalloc round trips (about a hundred buffers, 1 GB on the 119^3 lattice)
    for (auto &s : spare) lv.push_back(std::move(s));
    spare.swap(lv);
    lv.clear();
    rows.clear();
    nb = 0;
}

std::unique_ptr<MlCoarse> MlPrec::take_level()
{
    if (spare.empty()) return std::unique_ptr<MlCoarse>(new MlCoarse);
    std::unique_ptr<MlCoarse> c = std::move(spare.front());
    spare.erase(spare.begin());
    c->n = c->nnz = 0;
    c->has_next = c->smoothed = false;
    c->n_next = c->nnzp = c->nnzap = 0;
    return c;
}

void MlPrec::build_hierarchy_host(int64_t nbg, const int32_t *h_rowptr, const int32_t *h_colidx, const double *pos_rm,
                                  double mean_len, int world_, MlHierarchy &H, const FinePatternGate *fine_pattern) const
{
    std::vector<int64_t> bounds((size_t)world_ + 1);
    for (int q = 0; q <= world_; ++q) bounds[(size_t)q] = nbg * q / world_;
    build_ml_hierarchy(nbg, h_rowptr, h_colidx, pos_rm, par.cell * mean_len, par.ratio, par.max_levels, par.coarsest_rows,
                       std::max(1, par.smooth_from), H, world_ > 1 ? bounds.data() : nullptr, world_, fine_pattern);
}

bool MlPrec::build_symbolic(cudaStream_t st, int64_t nb_, const int32_t *h_rowptr, const int32_t *h_colidx, const double *pos_rm,
                            double mean_len, const MlDist *dist)
{
    if (nb_ <= 0 || !(mean_len > 0.0)) { reset(); return false; }
    const double t0w = wall_ms();
    MlHierarchy H;
    build_hierarchy_host(nb_, h_rowptr, h_colidx, pos_rm, mean_len, dist ? dist->world : 1, H);
    const bool ok = upload_hierarchy(st, H, nb_, h_rowptr[nb_], dist ? h_rowptr[dist->row_begin] : 0, dist);
    ms_symbolic = wall_ms() - t0w;
    return ok;
}

bool MlPrec::upload_hierarchy(cudaStream_t st, const MlHierarchy &H, int64_t nbg, int64_t fine_nnz, int64_t own_first_block,
                              const MlDist *dist)
{
    const double t0w = wall_ms();
    reset();
    if (nbg <= 0) return false;
    half16 = [] { const char *e = std::getenv("MBFEA_ML_HALF"); return e && *e == '1'; }();
    rank = dist ? dist->rank : 0; world = dist ? dist->world : 1;
    nc = dist ? dist->nc : nullptr; comm = dist ? dist->comm : nullptr;
    std::vector<int64_t> bounds((size_t)world + 1);
    for (int q = 0; q <= world; ++q) bounds[(size_t)q] = nbg * q / world;
    const size_t L = H.levels();
    if (L == 0) return false;
    if (6 * H.agg[L - 1].n_coarse > 3072) return false;   // the chain stopped early (max_levels): no dense inverse that large
    // level 0 -> 1
    const HierLevel &h0 = H.agg[0];
    int32_t biggest = 0;
    for (int64_t I = 0; I < h0.n_coarse; ++I) biggest = std::max(biggest, h0.agg_ptr[(size_t)I + 1] - h0.agg_ptr[(size_t)I]);
    group0 = biggest <= 8 ? 8 : 32;
    agg_begin.assign((size_t)world + 1, 0); blk_begin.assign((size_t)world + 1, 0);
    for (int q = 0; q < world; ++q) {   // aggregates are numbered rank-major: rank q owns [agg of its first row .. )
        int32_t hi = -1;
        for (int64_t v = bounds[(size_t)q]; v < bounds[(size_t)q + 1]; ++v) hi = std::max(hi, h0.agg[(size_t)v]);
        agg_begin[(size_t)q + 1] = std::max<int64_t>(agg_begin[(size_t)q], (int64_t)hi + 1);
    }
    for (int q = 0; q <= world; ++q) blk_begin[(size_t)q] = h0.rowptr[(size_t)agg_begin[(size_t)q]];
    if (world == 1) {
        nb = nbg;
        up(st, agg0, h0.agg); up(st, agg_ptr0, h0.agg_ptr); up(st, agg_rows0, h0.agg_rows);
        up(st, gal_ptr0, h0.gal_ptr); up(st, gal_src0, h0.gal_src); up(st, off0, h0.off);
        CK(frow0.alloc((size_t)fine_nnz));
    } else {
        // the rank's share of level 0: own rows (local ids), own aggregates, own rows of the level-1 matrix
        const int64_t r0 = dist->row_begin, nbl = dist->row_end - dist->row_begin;
        const int64_t ng = (int64_t)dist->ghost_global->size();
        const int64_t a0 = agg_begin[(size_t)rank], a1 = agg_begin[(size_t)rank + 1];
        nb = nbl;
        std::vector<int32_t> agg_l((size_t)nbl), ptr_l((size_t)(a1 - a0) + 1), rows_l;
        std::vector<double> off_l((size_t)3 * (size_t)(nbl + ng));
        for (int64_t i = 0; i < nbl; ++i) {
            agg_l[(size_t)i] = h0.agg[(size_t)(r0 + i)];
            for (int a = 0; a < 3; ++a) off_l[(size_t)3 * i + a] = h0.off[(size_t)3 * (r0 + i) + a];
        }
        for (int64_t g = 0; g < ng; ++g)
            for (int a = 0; a < 3; ++a) off_l[(size_t)3 * (nbl + g) + a] = h0.off[(size_t)3 * (*dist->ghost_global)[(size_t)g] + a];
        rows_l.reserve((size_t)nbl);
        for (int64_t I = a0; I < a1; ++I) {
            ptr_l[(size_t)(I - a0)] = (int32_t)rows_l.size();
            for (int32_t q = h0.agg_ptr[(size_t)I]; q < h0.agg_ptr[(size_t)I + 1]; ++q) rows_l.push_back((int32_t)(h0.agg_rows[(size_t)q] - r0));
        }
        ptr_l[(size_t)(a1 - a0)] = (int32_t)rows_l.size();
        // contributor lists of the own level-1 blocks, fine blocks renumbered from the global to the local pattern
        const int64_t b0 = blk_begin[(size_t)rank], b1 = blk_begin[(size_t)rank + 1];
        std::vector<int32_t> gp((size_t)(b1 - b0) + 1), gs;
        gs.reserve((size_t)(h0.gal_ptr[(size_t)b1] - h0.gal_ptr[(size_t)b0]));
        // the rank's local pattern keeps the blocks of a row in GLOBAL column order (build_plan only renames the column
        // ids), so the local index of global fine block k of an own row is k - (first block of the first own row)
        const int32_t k_lo = (int32_t)own_first_block;
        for (int64_t k = b0; k < b1; ++k) {
            gp[(size_t)(k - b0)] = (int32_t)gs.size();
            for (int32_t s2 = h0.gal_ptr[(size_t)k]; s2 < h0.gal_ptr[(size_t)k + 1]; ++s2) gs.push_back(h0.gal_src[(size_t)s2] - k_lo);
        }
        gp[(size_t)(b1 - b0)] = (int32_t)gs.size();
        up(st, agg0, agg_l); up(st, agg_ptr0, ptr_l); up(st, agg_rows0, rows_l);
        up(st, gal_ptr0, gp); up(st, gal_src0, gs); up(st, off0, off_l);
        CK(frow0.alloc((size_t)dist->loc_rowptr[nbl]));
    }
    CK(t0.alloc((size_t)nb * 36));
    if (half16) CK(t0h.alloc((size_t)nb * kTBlock16Bytes));
    rows.push_back(nb);
    for (size_t m = 0; m < L; ++m) {
        std::unique_ptr<MlCoarse> c = take_level();
        const hvec<int32_t> &rp = m == 0 ? h0.rowptr : H.xfer[m].crow;
        const hvec<int32_t> &ci = m == 0 ? h0.colidx : H.xfer[m].ccol;
        c->n = (int64_t)rp.size() - 1;
        c->nnz = (int64_t)ci.size();
        rows.push_back(c->n);
        up(st, c->rowptr, rp); up(st, c->colidx, ci);
        std::vector<int32_t> diag((size_t)c->n);
        for (int64_t I = 0; I < c->n; ++I)
            diag[(size_t)I] = (int32_t)(std::lower_bound(ci.begin() + rp[(size_t)I], ci.begin() + rp[(size_t)I + 1], (int32_t)I) - ci.begin());
        up(st, c->diag_idx, diag);
        if (m + 1 < L) {   // transfer from level m + 1 to level m + 2
            const HierLevel &h = H.agg[m + 1];
            const MlTransfer &T = H.xfer[m + 1];
            c->has_next = true; c->smoothed = T.smoothed; c->n_next = T.n_coarse;
            c->nnzp = (int64_t)T.pcol.size(); c->nnzap = (int64_t)T.apcol.size();
            up(st, c->agg, h.agg); up(st, c->off, h.off);
            up(st, c->prow, T.prow); up(st, c->pcol, T.pcol);
            up(st, c->rrow, T.rrow); up(st, c->rblk, T.rblk); up(st, c->rfine, T.rfine);
            up(st, c->aprow, T.aprow); up(st, c->apcol, T.apcol);
        }
        alloc_level_buffers(st, *c, m + 1 < L);
        lv.push_back(std::move(c));
    }
    n_dense = 6 * lv.back()->n;
    CK(dense.alloc((size_t)(n_dense * n_dense)));
    CK(dense32.alloc((size_t)(n_dense * n_dense)));
    CK(gj_col.alloc((size_t)(n_dense + 24) * 24));
    CK(cudaStreamSynchronize(st));   // the host vectors above go out of scope
    symbolic_ok = true;
    ms_symbolic = wall_ms() - t0w;
    return true;
}

// ---- the symbolic half on the device (one rank) --------------------------------------------------------------
namespace {
template <class T>
T fetch(cudaStream_t st, const T *d)   // one scalar back from the device (synchronises the stream)
{
    T v;
    CK(cudaMemcpyAsync(&v, d, sizeof(T), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return v;
}
}  // namespace

bool MlPrec::build_hierarchy_device(cudaStream_t st, int64_t V, const Node4 *d_node4, const int32_t *d_fm, int64_t nb_,
                                    const int32_t *d_rowptr, const int32_t *d_colidx, int64_t fine_nnz, double mean_len, int nparts)
{
    const double t0w = wall_ms();
    reset();
    if (nb_ <= 0 || !(mean_len > 0.0) || !(par.cell > 0.0) || !(par.ratio > 1.0)) return false;
    half16 = [] { const char *e = std::getenv("MBFEA_ML_HALF"); return e && *e == '1'; }();
    rank = 0; world = 1; nc = nullptr; comm = nullptr;
    nb = nb_;
    CK(hd_pos[0].alloc((size_t)3 * nb)); CK(hd_pos[1].alloc((size_t)3 * nb));
    CK(hd_part.alloc(6 * 256)); CK(hd_box.alloc(6));
    CK(hd_key.alloc((size_t)nb)); CK(hd_len.alloc((size_t)nb + 2)); CK(hd_small.alloc(8));
    CK(hd_scan.alloc((size_t)scan_i32_scratch(std::max<int64_t>(fine_nnz, std::min<int64_t>(65 * nb + 2048, 268435456 + 8)) + 2)));
    launch_h_pos(st, V, d_node4, d_fm, hd_pos[0].p);
    const double *pos = hd_pos[0].p;
    int cur_pos = 0;
    const int32_t *crp = d_rowptr, *cci = d_colidx;
    double cell = par.cell * mean_len;
    int64_t n = nb;
    const int smooth_from = std::max(1, par.smooth_from);
    bool ok = true;
    int64_t nk0 = 0, na0 = 0;
    // experiment: a different brick growth between the coarse levels (fewer, smaller levels in the replicated tail)
    const double ratio_coarse = [] { const char *e = std::getenv("MBFEA_ML_RATIO_COARSE"); return e ? std::atof(e) : 0.0; }();
    const bool trace = std::getenv("MBFEA_PREP_TRACE") != nullptr;
    auto give_up = [&](const char *why, int l, int64_t a, int64_t b) {
        if (trace) std::fprintf(stderr, "[mbfea hierarchy] device build gives up at level %d: %s (%lld, %lld)\n", l, why, (long long)a, (long long)b);
        ok = false;
    };
    for (int l = 0; l + 1 < par.max_levels && n > par.coarsest_rows; ++l) {
        // ---- aggregates of level l: bricks of edge `cell` over the bounding box ----
        launch_h_box(st, n, pos, hd_part.p, hd_box.p);
        double box[6];
        CK(cudaMemcpyAsync(box, hd_box.p, sizeof box, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int64_t cnt[3];
        const int parts = (l == 0 && nparts > 1) ? nparts : 1;   // several ranks: fine-level aggregates never straddle a rank
        double nk_d = (double)parts;
        for (int a = 0; a < 3; ++a) {
            cnt[a] = std::max<int64_t>(1, (int64_t)std::floor((box[3 + a] - box[a]) / cell + 1e-9) + 1);
            nk_d *= (double)cnt[a];
        }
        int64_t bounds[kMaxWorld + 1] = {0};
        if (parts > 1) { if (parts > kMaxWorld) { give_up("too many parts", l, parts, 0); break; } for (int q = 0; q <= parts; ++q) bounds[q] = nb * q / parts; }
        // (the host builder switches from its counting sort to a comparison sort above 8 n bricks -- same numbering either
        //  way; here the counting arrays may be far sparser before they stop paying: 64 n bricks, at most 2^28)
        if (!(nk_d <= 64.0 * (double)n + 1024.0) || nk_d > 268435456.0) { give_up("sparse brick grid", l, (int64_t)nk_d, n); break; }
        const int64_t nk = cnt[0] * cnt[1] * cnt[2] * parts;
        if (l == 0) {   // the brick grids of the coarser levels are smaller: one allocation serves every level
            nk0 = nk;
            CK(hd_bcnt.alloc((size_t)nk + 3)); CK(hd_bcur.alloc((size_t)nk + 3)); CK(hd_occ.alloc((size_t)nk + 3));
            CK(hd_aggid.alloc((size_t)nk + 3)); CK(hd_bstart.alloc((size_t)nk + 3));
        } else if (nk > nk0) { give_up("brick grid grew", l, nk, nk0); break; }
        CK(cudaMemsetAsync(hd_bcnt.p, 0, ((size_t)nk + 1) * sizeof(int32_t), st));
        CK(cudaMemsetAsync(hd_bcur.p, 0, ((size_t)nk + 1) * sizeof(int32_t), st));
        launch_h_brick(st, n, pos, box, cell, cnt, hd_key.p, hd_bcnt.p, bounds, parts);
        launch_h_occ(st, nk, hd_bcnt.p, hd_occ.p);
        launch_scan_i32(st, nk, hd_occ.p, hd_aggid.p, hd_scan.p);
        launch_scan_i32(st, nk, hd_bcnt.p, hd_bstart.p, hd_scan.p);
        const int64_t na = fetch(st, hd_aggid.p + nk);
        if (na >= n) break;   // no coarsening at this cell size: stop
        MlCoarse *cur = l >= 1 ? lv[(size_t)l - 1].get() : nullptr;   // level l (l >= 1) and its transfer to level l + 1
        DevBuf<int32_t> &agg = l == 0 ? agg0 : cur->agg;
        DevBuf<double> &off = l == 0 ? off0 : cur->off;
        DevBuf<int32_t> &aptr = l == 0 ? agg_ptr0 : hd_aggptr, &arows = l == 0 ? agg_rows0 : hd_aggrows;
        CK(agg.alloc((size_t)n)); CK(off.alloc((size_t)3 * n));
        if (l == 0) {
            na0 = na;
            CK(agg_ptr0.alloc((size_t)na + 1)); CK(agg_rows0.alloc((size_t)n));
            CK(hd_aggptr.alloc((size_t)na + 2)); CK(hd_aggrows.alloc((size_t)na + 2));   // member lists of the coarser levels (n <= na0)
        } else if (na + 1 > na0 + 2 || n > na0 + 2) { give_up("member scratch", l, na, n); break; }
        double *centroid = hd_pos[cur_pos ^ 1].p;
        launch_h_assign(st, n, hd_key.p, hd_aggid.p, hd_bstart.p, hd_bcur.p, agg.p, arows.p);
        launch_h_aggptr(st, nk, hd_bcnt.p, hd_aggid.p, hd_bstart.p, aptr.p, n, na);
        CK(cudaMemsetAsync(hd_small.p, 0, 8 * sizeof(int32_t), st));
        launch_h_sort_members(st, na, aptr.p, arows.p, hd_small.p);
        launch_h_centroid_off(st, n, na, aptr.p, arows.p, agg.p, pos, centroid, off.p);
        std::unique_ptr<MlCoarse> next = take_level();
        next->n = na;
        CK(next->rowptr.alloc((size_t)na + 1)); CK(next->diag_idx.alloc((size_t)na));
        if (l == 0) {
            group0 = fetch(st, hd_small.p) <= 8 ? 8 : 32;
            agg_begin.assign((size_t)parts + 1, 0);
            agg_begin[(size_t)parts] = na;
            for (int q = 1; q < parts; ++q) agg_begin[(size_t)q] = fetch(st, hd_aggid.p + (int64_t)q * (nk / parts));   // occupied bricks before part q
            // level-1 pattern: candidates per aggregate -> sorted unique
            launch_h0_ub(st, na, aptr.p, arows.p, crp, hd_len.p);
            // (hd_aggptr is free on level 0: start of every aggregate's candidate segment)
            launch_scan_i32(st, na, hd_len.p, hd_aggptr.p, hd_scan.p);
            const int64_t ncand = fetch(st, hd_aggptr.p + na);
            CK(hd_cand.alloc((size_t)ncand + 1));
            launch_h0_cand(st, na, aptr.p, arows.p, crp, cci, agg.p, hd_aggptr.p, hd_cand.p, hd_len.p);
            launch_scan_i32(st, na, hd_len.p, next->rowptr.p, hd_scan.p);
            next->nnz = fetch(st, next->rowptr.p + na);
            CK(next->colidx.alloc((size_t)next->nnz));
            launch_h0_emit(st, na, hd_aggptr.p, hd_cand.p, next->rowptr.p, next->colidx.p, next->diag_idx.p);
            blk_begin.assign((size_t)parts + 1, 0);
            blk_begin[(size_t)parts] = next->nnz;
            for (int q = 1; q < parts; ++q) blk_begin[(size_t)q] = fetch(st, next->rowptr.p + agg_begin[(size_t)q]);
            // Galerkin contributor lists (counts, prefix, fill; the counters live in the candidate scratch)
            CK(hd_cand.alloc((size_t)std::max<int64_t>(ncand, next->nnz) + 1));
            CK(gal_ptr0.alloc((size_t)next->nnz + 1)); CK(gal_src0.alloc((size_t)fine_nnz));
            CK(cudaMemsetAsync(hd_cand.p, 0, ((size_t)next->nnz + 1) * sizeof(int32_t), st));
            launch_h0_gal(st, false, na, aptr.p, arows.p, crp, cci, agg.p, next->rowptr.p, next->colidx.p, hd_cand.p, nullptr, nullptr);
            launch_scan_i32(st, next->nnz, hd_cand.p, gal_ptr0.p, hd_scan.p);
            CK(cudaMemsetAsync(hd_cand.p, 0, ((size_t)next->nnz + 1) * sizeof(int32_t), st));
            launch_h0_gal(st, true, na, aptr.p, arows.p, crp, cci, agg.p, next->rowptr.p, next->colidx.p, hd_cand.p, gal_ptr0.p, gal_src0.p);
            CK(frow0.alloc((size_t)fine_nnz));
            CK(t0.alloc((size_t)nb * 36));
            if (half16) CK(t0h.alloc((size_t)nb * kTBlock16Bytes));
            rows.push_back(nb);
        } else {
            // transfer from level l to level l + 1: P, P^T, A P, P^T A P (patterns)
            cur->has_next = true; cur->smoothed = l >= smooth_from; cur->n_next = na;
            if (na > 200 * 1024 * 8) { give_up("row bitmap beyond shared memory", l, na, 0); break; }   // host builder
            CK(cur->prow.alloc((size_t)n + 1));
            launch_hx_p_rows(st, false, n, na, cur->smoothed, crp, cci, agg.p, hd_len.p, nullptr, nullptr);
            launch_scan_i32(st, n, hd_len.p, cur->prow.p, hd_scan.p);
            cur->nnzp = fetch(st, cur->prow.p + n);
            CK(cur->pcol.alloc((size_t)cur->nnzp));
            launch_hx_p_rows(st, true, n, na, cur->smoothed, crp, cci, agg.p, nullptr, cur->prow.p, cur->pcol.p);
            CK(cur->prow_of.alloc((size_t)cur->nnzp));
            launch_ml_row_of(st, n, cur->prow.p, cur->prow_of.p);
            // P^T
            CK(cur->rrow.alloc((size_t)na + 1)); CK(cur->rblk.alloc((size_t)cur->nnzp)); CK(cur->rfine.alloc((size_t)cur->nnzp));
            int32_t *tcnt = hd_bcur.p;   // free again after the member fill (na <= nk)
            CK(cudaMemsetAsync(tcnt, 0, ((size_t)na + 2) * sizeof(int32_t), st));
            launch_hx_transpose(st, 0, cur->nnzp, na, cur->pcol.p, nullptr, tcnt, nullptr, nullptr, nullptr);
            launch_scan_i32(st, na, tcnt, cur->rrow.p, hd_scan.p);
            CK(cudaMemsetAsync(tcnt, 0, ((size_t)na + 2) * sizeof(int32_t), st));
            launch_hx_transpose(st, 1, cur->nnzp, na, cur->pcol.p, nullptr, tcnt, cur->rrow.p, cur->rblk.p, nullptr);
            launch_hx_transpose(st, 2, cur->nnzp, na, nullptr, cur->prow_of.p, nullptr, cur->rrow.p, cur->rblk.p, cur->rfine.p);
            // A P
            CK(cur->aprow.alloc((size_t)n + 1));
            launch_hx_ap_rows(st, false, n, na, crp, cci, cur->prow.p, cur->pcol.p, hd_len.p, nullptr, nullptr);
            launch_scan_i32(st, n, hd_len.p, cur->aprow.p, hd_scan.p);
            cur->nnzap = fetch(st, cur->aprow.p + n);
            CK(cur->apcol.alloc((size_t)cur->nnzap));
            launch_hx_ap_rows(st, true, n, na, crp, cci, cur->prow.p, cur->pcol.p, nullptr, cur->aprow.p, cur->apcol.p);
            // P^T (A P): the next level's matrix
            launch_hx_rap_rows(st, false, na, cur->rrow.p, cur->rfine.p, cur->aprow.p, cur->apcol.p, hd_len.p, nullptr, nullptr);
            launch_scan_i32(st, na, hd_len.p, next->rowptr.p, hd_scan.p);
            next->nnz = fetch(st, next->rowptr.p + na);
            CK(next->colidx.alloc((size_t)next->nnz));
            launch_hx_rap_rows(st, true, na, cur->rrow.p, cur->rfine.p, cur->aprow.p, cur->apcol.p, nullptr, next->rowptr.p, next->colidx.p);
            launch_h_diag(st, na, next->rowptr.p, next->colidx.p, next->diag_idx.p);
        }
        rows.push_back(na);
        crp = next->rowptr.p; cci = next->colidx.p;
        lv.push_back(std::move(next));
        cur_pos ^= 1;
        pos = hd_pos[cur_pos].p;
        n = na;
        cell *= (l >= 1 && ratio_coarse > 1.0) ? ratio_coarse : par.ratio;
    }
    if (ok && !lv.empty() && 6 * lv.back()->n > 3072) give_up("coarsest level too large", (int)lv.size(), lv.back()->n, 0);
    if (!ok || lv.empty()) { reset(); return false; }
    const size_t L = lv.size();
    for (size_t m = 0; m < L; ++m) alloc_level_buffers(st, *lv[m], m + 1 < L);
    n_dense = 6 * lv.back()->n;
    CK(dense.alloc((size_t)(n_dense * n_dense)));
    CK(dense32.alloc((size_t)(n_dense * n_dense)));
    CK(gj_col.alloc((size_t)(n_dense + 24) * 24));
    CK(cudaStreamSynchronize(st));
    symbolic_ok = true;
    ms_symbolic = wall_ms() - t0w;
    return true;
}

bool MlPrec::pack_image_device(cudaStream_t st, int world_, int64_t nbg, const int32_t *d_fine_rowptr, int64_t fine_nnz,
                               DevBuf<unsigned char> &dimg, MlImageHeader &hd)
{
    const size_t L = lv.size();
    if (!symbolic_ok || L == 0 || L > (size_t)kMlImageMaxLevels || world_ < 1 || world_ > kMlImageMaxWorld) return false;
    if ((int)agg_begin.size() != world_ + 1 || (int)blk_begin.size() != world_ + 1) return false;
    std::memset(&hd, 0, sizeof hd);
    hd.magic = kMlImageMagic; hd.levels = (int64_t)L; hd.world = world_; hd.nbg = nbg; hd.fine_nnz = fine_nnz; hd.group0 = group0;
    for (int q = 0; q <= world_; ++q) {
        hd.agg_begin[q] = agg_begin[(size_t)q];
        hd.blk_begin[q] = blk_begin[(size_t)q];
        hd.own_first[q] = fetch(st, d_fine_rowptr + nbg * q / world_);
        hd.aggptr_at[q] = fetch(st, agg_ptr0.p + agg_begin[(size_t)q]);
        hd.galptr_at[q] = fetch(st, gal_ptr0.p + blk_begin[(size_t)q]);
    }
    struct Part { const void *p; size_t bytes; int64_t cnt; };
    std::vector<std::vector<Part>> src(L);
    auto pi = [](const DevBuf<int32_t> &d, int64_t n) { return Part{d.p, (size_t)n * sizeof(int32_t), n}; };
    auto pd = [](const DevBuf<double> &d, int64_t n) { return Part{d.p, (size_t)n * sizeof(double), n}; };
    const MlCoarse &c0 = *lv[0];
    src[0] = {pi(agg0, nb), pd(off0, 3 * nb), pi(agg_ptr0, c0.n + 1), pi(agg_rows0, nb), pi(c0.rowptr, c0.n + 1), pi(c0.colidx, c0.nnz),
              pi(c0.diag_idx, c0.n), pi(gal_ptr0, c0.nnz + 1), pi(gal_src0, fine_nnz)};
    for (size_t l = 1; l < L; ++l) {
        const MlCoarse &c = *lv[l - 1], &nx = *lv[l];
        src[l] = {pi(c.agg, c.n), pd(c.off, 3 * c.n), pi(c.prow, c.n + 1), pi(c.pcol, c.nnzp), pi(c.rrow, nx.n + 1), pi(c.rblk, c.nnzp),
                  pi(c.rfine, c.nnzp), pi(c.aprow, c.n + 1), pi(c.apcol, c.nnzap), pi(nx.rowptr, nx.n + 1), pi(nx.colidx, nx.nnz),
                  pi(nx.diag_idx, nx.n)};
    }
    size_t at = (sizeof(MlImageHeader) + 15) & ~(size_t)15;
    for (size_t l = 0; l < L; ++l) {
        hd.lv[l].n_fine = l == 0 ? nb : lv[l - 1]->n; hd.lv[l].n_coarse = lv[l]->n;
        hd.lv[l].smoothed = (l >= 1 && lv[l - 1]->smoothed) ? 1 : 0;
        for (size_t k = 0; k < src[l].size(); ++k) {
            hd.lv[l].off[k] = (int64_t)at; hd.lv[l].cnt[k] = src[l][k].cnt;
            at += (src[l][k].bytes + 15) & ~(size_t)15;
        }
    }
    hd.total_bytes = (int64_t)at;
    CK(dimg.alloc(at));
    CK(cudaMemcpyAsync(dimg.p, &hd, sizeof hd, cudaMemcpyHostToDevice, st));
    for (size_t l = 0; l < L; ++l)
        for (size_t k = 0; k < src[l].size(); ++k)
            if (src[l][k].bytes)
                CK(cudaMemcpyAsync(dimg.p + hd.lv[l].off[k], src[l][k].p, src[l][k].bytes, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));   // (hd is the caller's: the header copy above has been staged)
    return true;
}

// parity tap: device arrays of the current hierarchy against a host hierarchy.  bad: per level 16 counters --
// level 0: agg, off, agg_ptr, agg_rows, rowptr1, colidx1, diag1, gal_ptr, gal_src; level l >= 1: agg, off, prow, pcol,
// rrow, rblk, rfine, aprow, apcol, crow, ccol, cdiag  (-1: different length, -2: level missing)
void MlPrec::compare_with_host(cudaStream_t st, const MlHierarchy &H, int64_t fine_nnz, std::vector<int64_t> &bad) const
{
    CK(cudaStreamSynchronize(st));
    const size_t L = std::max(H.levels(), lv.size());
    bad.assign(16 * std::max<size_t>(L, 1), 0);
    auto cmp_i = [&](const DevBuf<int32_t> &d, size_t n_dev, const auto &h) -> int64_t {
        if (n_dev != h.size()) return -1;
        std::vector<int32_t> t(h.size());
        if (!t.empty()) CK(cudaMemcpy(t.data(), d.p, t.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        int64_t b = 0;
        for (size_t i = 0; i < t.size(); ++i) b += t[i] != h[i];
        return b;
    };
    auto cmp_d = [&](const DevBuf<double> &d, size_t n_dev, const auto &h) -> int64_t {
        if (n_dev != h.size()) return -1;
        std::vector<double> t(h.size());
        if (!t.empty()) CK(cudaMemcpy(t.data(), d.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost));
        int64_t b = 0;
        for (size_t i = 0; i < t.size(); ++i) b += !(t[i] == h[i]);
        return b;
    };
    if (H.levels() != lv.size()) { for (auto &x : bad) x = -2; return; }
    if (L == 0) return;
    const HierLevel &h0 = H.agg[0];
    const MlCoarse &c0 = *lv[0];
    std::vector<int32_t> diag1((size_t)h0.n_coarse);
    for (int64_t I = 0; I < h0.n_coarse; ++I)
        diag1[(size_t)I] = (int32_t)(std::lower_bound(h0.colidx.begin() + h0.rowptr[(size_t)I], h0.colidx.begin() + h0.rowptr[(size_t)I + 1],
                                                      (int32_t)I) - h0.colidx.begin());
    bad[0] = cmp_i(agg0, (size_t)nb, h0.agg); bad[1] = cmp_d(off0, (size_t)3 * nb, h0.off);
    bad[2] = cmp_i(agg_ptr0, (size_t)c0.n + 1, h0.agg_ptr); bad[3] = cmp_i(agg_rows0, (size_t)nb, h0.agg_rows);
    bad[4] = cmp_i(c0.rowptr, (size_t)c0.n + 1, h0.rowptr); bad[5] = cmp_i(c0.colidx, (size_t)c0.nnz, h0.colidx);
    bad[6] = cmp_i(c0.diag_idx, (size_t)c0.n, diag1);
    bad[7] = cmp_i(gal_ptr0, (size_t)c0.nnz + 1, h0.gal_ptr); bad[8] = cmp_i(gal_src0, (size_t)fine_nnz, h0.gal_src);
    for (size_t l = 1; l < L; ++l) {
        const HierLevel &a = H.agg[l];
        const MlTransfer &T = H.xfer[l];
        const MlCoarse &c = *lv[l - 1], &nx = *lv[l];
        int64_t *b = bad.data() + 16 * l;
        b[0] = cmp_i(c.agg, (size_t)c.n, a.agg); b[1] = cmp_d(c.off, (size_t)3 * c.n, a.off);
        b[2] = cmp_i(c.prow, (size_t)c.n + 1, T.prow); b[3] = cmp_i(c.pcol, (size_t)c.nnzp, T.pcol);
        b[4] = cmp_i(c.rrow, (size_t)nx.n + 1, T.rrow); b[5] = cmp_i(c.rblk, (size_t)c.nnzp, T.rblk); b[6] = cmp_i(c.rfine, (size_t)c.nnzp, T.rfine);
        b[7] = cmp_i(c.aprow, (size_t)c.n + 1, T.aprow); b[8] = cmp_i(c.apcol, (size_t)c.nnzap, T.apcol);
        b[9] = cmp_i(nx.rowptr, (size_t)nx.n + 1, T.crow); b[10] = cmp_i(nx.colidx, (size_t)nx.nnz, T.ccol); b[11] = cmp_i(nx.diag_idx, (size_t)nx.n, T.cdiag);
        b[12] = (c.smoothed == T.smoothed) ? 0 : 1;
    }
}

// value / work arrays of a coarse level whose index arrays are in place
void MlPrec::alloc_level_buffers(cudaStream_t st, MlCoarse &c, bool has_next)
{
    CK(c.row_of.alloc((size_t)c.nnz));
    CK(c.vals.alloc((size_t)c.nnz * 36));
    CK(c.r.alloc((size_t)c.n * 6)); CK(c.z.alloc((size_t)c.n * 6));
    launch_ml_row_of(st, c.n, c.rowptr.p, c.row_of.p);
    if (!has_next) return;
    CK(c.prow_of.alloc((size_t)c.nnzp)); CK(c.aprow_of.alloc((size_t)c.nnzap));
    launch_ml_row_of(st, c.n, c.prow.p, c.prow_of.p);
    launch_ml_row_of(st, c.n, c.aprow.p, c.aprow_of.p);
    CK(c.pval.alloc((size_t)c.nnzp * 36)); CK(c.apval.alloc((size_t)c.nnzap * 36));
    if (c.smoothed) { if (half16) CK(c.pvalh.alloc((size_t)c.nnzp * kTBlock16Bytes)); else CK(c.pval32.alloc((size_t)c.nnzp * 36)); }
    CK(c.lfac.alloc((size_t)c.n * 36)); CK(c.sinv.alloc((size_t)c.n * 36)); CK(c.dinv.alloc((size_t)c.n * 36));
    CK(c.dinv32.alloc((size_t)c.n * 36));
}

bool MlPrec::upload_from_image(cudaStream_t st, const unsigned char *d_img, const MlImageHeader &hd, const MlDist &dist)
{
    const double t0w = wall_ms();
    reset();
    if (hd.magic != kMlImageMagic || hd.levels < 1 || hd.levels > kMlImageMaxLevels || hd.world != dist.world) return false;
    half16 = [] { const char *e = std::getenv("MBFEA_ML_HALF"); return e && *e == '1'; }();
    rank = dist.rank; world = dist.world; nc = dist.nc; comm = dist.comm;
    const size_t L = (size_t)hd.levels;
    if (6 * hd.lv[L - 1].n_coarse > 3072) return false;
    group0 = (int)hd.group0;
    agg_begin.assign(hd.agg_begin, hd.agg_begin + world + 1);
    blk_begin.assign(hd.blk_begin, hd.blk_begin + world + 1);
    auto src_i = [&](size_t l, int k) { return reinterpret_cast<const int32_t *>(d_img + hd.lv[l].off[k]); };
    auto src_d = [&](size_t l, int k) { return reinterpret_cast<const double *>(d_img + hd.lv[l].off[k]); };
    auto copy_i = [&](DevBuf<int32_t> &dst, size_t l, int k) {
        CK(dst.alloc((size_t)hd.lv[l].cnt[k]));
        if (hd.lv[l].cnt[k]) CK(cudaMemcpyAsync(dst.p, src_i(l, k), (size_t)hd.lv[l].cnt[k] * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    };
    auto copy_d = [&](DevBuf<double> &dst, size_t l, int k) {
        CK(dst.alloc((size_t)hd.lv[l].cnt[k]));
        if (hd.lv[l].cnt[k]) CK(cudaMemcpyAsync(dst.p, src_d(l, k), (size_t)hd.lv[l].cnt[k] * sizeof(double), cudaMemcpyDeviceToDevice, st));
    };
    // ---- level 0: the rank's own rows, aggregates and level-1 block rows
    const int64_t r0 = dist.row_begin, nbl = dist.row_end - dist.row_begin, ng = (int64_t)dist.ghost_global->size();
    const int64_t a0 = agg_begin[(size_t)rank], a1 = agg_begin[(size_t)rank + 1];
    const int64_t b0 = blk_begin[(size_t)rank], b1 = blk_begin[(size_t)rank + 1];
    const int64_t m0 = hd.aggptr_at[rank], m1 = hd.aggptr_at[rank + 1];   // member-list range of the own aggregates
    const int64_t g0 = hd.galptr_at[rank], g1 = hd.galptr_at[rank + 1];   // contributor-list range of the own level-1 blocks
    nb = nbl;
    CK(agg0.alloc((size_t)nbl));
    CK(cudaMemcpyAsync(agg0.p, src_i(0, 0) + r0, (size_t)nbl * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    CK(off0.alloc((size_t)3 * (size_t)(nbl + ng)));
    CK(cudaMemcpyAsync(off0.p, src_d(0, 1) + 3 * r0, (size_t)3 * nbl * sizeof(double), cudaMemcpyDeviceToDevice, st));
    DevBuf<int64_t> &d_ghost = ghost_stage;   // kept between jobs
    if (ng > 0) {
        CK(d_ghost.alloc((size_t)ng));
        CK(cudaMemcpyAsync(d_ghost.p, dist.ghost_global->data(), (size_t)ng * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        launch_ml_gather3(st, ng, src_d(0, 1), d_ghost.p, off0.p + 3 * nbl);
    }
    CK(agg_ptr0.alloc((size_t)(a1 - a0) + 1));
    launch_ml_shift_i32(st, a1 - a0 + 1, src_i(0, 2) + a0, (int32_t)m0, agg_ptr0.p);
    CK(agg_rows0.alloc((size_t)(m1 - m0)));
    launch_ml_shift_i32(st, m1 - m0, src_i(0, 3) + m0, (int32_t)r0, agg_rows0.p);
    CK(gal_ptr0.alloc((size_t)(b1 - b0) + 1));
    launch_ml_shift_i32(st, b1 - b0 + 1, src_i(0, 7) + b0, (int32_t)g0, gal_ptr0.p);
    CK(gal_src0.alloc((size_t)(g1 - g0)));
    // the rank's local pattern keeps the blocks of a row in GLOBAL column order, so global fine block k of an own row is
    // local block k - (first block of the rank's first row)
    launch_ml_shift_i32(st, g1 - g0, src_i(0, 8) + g0, (int32_t)hd.own_first[rank], gal_src0.p);
    CK(frow0.alloc((size_t)dist.loc_rowptr[nbl]));
    CK(t0.alloc((size_t)nb * 36));
    if (half16) CK(t0h.alloc((size_t)nb * kTBlock16Bytes));
    rows.push_back(nb);
    // ---- coarse levels: replicated
    for (size_t m = 0; m < L; ++m) {
        std::unique_ptr<MlCoarse> c = take_level();
        c->n = hd.lv[m].n_coarse;
        if (m == 0) { copy_i(c->rowptr, 0, 4); copy_i(c->colidx, 0, 5); copy_i(c->diag_idx, 0, 6); c->nnz = hd.lv[0].cnt[5]; }
        else { copy_i(c->rowptr, m, 9); copy_i(c->colidx, m, 10); copy_i(c->diag_idx, m, 11); c->nnz = hd.lv[m].cnt[10]; }
        rows.push_back(c->n);
        const bool has_next = m + 1 < L;
        if (has_next) {   // transfer from level m + 1 to level m + 2: image level m + 1
            const size_t l = m + 1;
            c->has_next = true; c->smoothed = hd.lv[l].smoothed != 0; c->n_next = hd.lv[l].n_coarse;
            c->nnzp = hd.lv[l].cnt[3]; c->nnzap = hd.lv[l].cnt[8];
            copy_i(c->agg, l, 0); copy_d(c->off, l, 1); copy_i(c->prow, l, 2); copy_i(c->pcol, l, 3);
            copy_i(c->rrow, l, 4); copy_i(c->rblk, l, 5); copy_i(c->rfine, l, 6); copy_i(c->aprow, l, 7); copy_i(c->apcol, l, 8);
        }
        alloc_level_buffers(st, *c, has_next);
        lv.push_back(std::move(c));
    }
    n_dense = 6 * lv.back()->n;
    CK(dense.alloc((size_t)(n_dense * n_dense)));
    CK(dense32.alloc((size_t)(n_dense * n_dense)));
    CK(gj_col.alloc((size_t)(n_dense + 24) * 24));
    CK(cudaStreamSynchronize(st));   // d_ghost and the image may go away
    symbolic_ok = true;
    ms_symbolic = wall_ms() - t0w;
    return true;
}

void MlPrec::build_numeric(cudaStream_t st, const int32_t *d_rowptr, const int32_t *d_colidx, const double *d_vals,
                           const double *d_lfac, int *d_flag)
{
    numeric_ok = false;
    if (!symbolic_ok) return;
    int64_t k = 0;
    launch_ml_row_of(st, nb, d_rowptr, frow0.p);
    launch_ml_make_t0(st, nb, d_lfac, off0.p, t0.p);
    if (half16) { launch_ml_pack16(st, nb, t0.p, t0h.p); ++k; }
    // level-1 matrix: the own block rows (all of them on one rank), then the slices of the other ranks
    launch_ml_galerkin0(st, blk_begin[(size_t)rank + 1] - blk_begin[(size_t)rank], gal_ptr0.p, gal_src0.p, frow0.p, d_colidx, d_vals,
                        off0.p, lv[0]->vals.p + 36 * blk_begin[(size_t)rank]);
    k += 3;
    if (world > 1) { exchange_slices(st, lv[0]->vals.p, blk_begin, 36); ++k; }
    if (std::getenv("MBFEA_ML_DEBUG")) {   // level-1 matrix after the exchange: zero blocks and asymmetry per owner rank
        MlCoarse &c1 = *lv[0];
        std::vector<double> v((size_t)c1.nnz * 36);
        std::vector<int32_t> rp((size_t)c1.n + 1), ci((size_t)c1.nnz);
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(v.data(), c1.vals.p, v.size() * sizeof(double), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rp.data(), c1.rowptr.p, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ci.data(), c1.colidx.p, ci.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int q = 0; q < world; ++q) {
            int64_t zero = 0, nonfinite = 0, badsym = 0, neg_diag = 0;
            double worst = 0.0;
            for (int64_t I = agg_begin[(size_t)q]; I < agg_begin[(size_t)q + 1]; ++I)
                for (int32_t kk = rp[(size_t)I]; kk < rp[(size_t)I + 1]; ++kk) {
                    const double *B = v.data() + (size_t)kk * 36;
                    double mx = 0.0; bool fin = true;
                    for (int e = 0; e < 36; ++e) { mx = std::max(mx, std::fabs(B[e])); fin = fin && std::isfinite(B[e]); }
                    if (!fin) { ++nonfinite; continue; }
                    if (mx == 0.0) ++zero;
                    const int32_t J = ci[(size_t)kk];
                    if (J == I) { for (int d = 0; d < 6; ++d) if (!(B[7 * d] > 0.0)) { ++neg_diag; break; } }
                    const int32_t *jb = ci.data() + rp[(size_t)J], *je = ci.data() + rp[(size_t)J + 1];
                    const int32_t *pos = std::lower_bound(jb, je, (int32_t)I);
                    if (pos == je || *pos != I) { ++badsym; continue; }
                    const double *T = v.data() + (size_t)(pos - ci.data()) * 36;
                    for (int r = 0; r < 6; ++r) for (int cc = 0; cc < 6; ++cc) worst = std::max(worst, std::fabs(B[6 * r + cc] - T[6 * cc + r]) / (mx + 1e-300));
                }
            std::fprintf(stderr, "[mbfea ml debug] rank %d sees level-1 rows of rank %d: aggregates [%lld, %lld) blocks [%lld, %lld): zero blocks %lld, non-finite %lld, "
                                 "missing transposes %lld, non-positive diagonals %lld, worst asymmetry %.2e\n", rank, q,
                         (long long)agg_begin[(size_t)q], (long long)agg_begin[(size_t)q + 1], (long long)blk_begin[(size_t)q],
                         (long long)blk_begin[(size_t)q + 1], (long long)zero, (long long)nonfinite, (long long)badsym, (long long)neg_diag, worst);
        }
    }
    for (size_t m = 0; m < lv.size(); ++m) {
        MlCoarse &c = *lv[m];
        if (!c.has_next) break;
        MlCoarse &nx = *lv[m + 1];
        launch_block_chol(st, c.n, c.diag_idx.p, c.vals.p, c.lfac.p, c.sinv.p, d_flag);
        launch_ml_diag_inv(st, c.n, c.sinv.p, c.dinv.p, c.dinv32.p);
        launch_ml_make_p(st, c.n, c.nnzp, c.prow.p, c.pcol.p, c.prow_of.p, c.rowptr.p, c.colidx.p, c.vals.p, c.agg.p, c.off.p,
                         c.dinv.p, c.smoothed ? par.omega : 0.0, c.pval.p);
        if (c.smoothed) {
            if (half16) launch_ml_pack16(st, c.nnzp, c.pval.p, c.pvalh.p);
            else launch_ml_to_f32(st, c.nnzp * 36, c.pval.p, c.pval32.p);
            ++k;
        }
        launch_ml_ap(st, c.n, c.nnzap, c.aprow.p, c.apcol.p, c.aprow_of.p, c.rowptr.p, c.colidx.p, c.vals.p, c.prow.p, c.pcol.p,
                     c.pval.p, c.apval.p);
        launch_ml_rap(st, nx.n, nx.nnz, nx.rowptr.p, nx.colidx.p, nx.row_of.p, c.rrow.p, c.rblk.p, c.rfine.p, c.pval.p, c.aprow.p,
                      c.apcol.p, c.apval.p, nx.vals.p);
        k += 5;
    }
    MlCoarse &last = *lv.back();
    CK(cudaMemsetAsync(dense.p, 0, (size_t)(n_dense * n_dense) * sizeof(double), st));
    launch_ml_blocks_to_dense(st, last.n, last.nnz, last.rowptr.p, last.colidx.p, last.row_of.p, last.vals.p, dense.p);
    launch_ml_dense_inverse(st, (int)n_dense, dense.p, gj_col.p, d_flag);
    launch_ml_dense_to_f32(st, (int)n_dense, dense.p, dense32.p);
    k += 2 + 3 * ((n_dense + 23) / 24);
    setup_launches = k;
    numeric_ok = true;
}

// every rank sends its slice [begin[rank], begin[rank + 1]) (x width doubles) of `buf` to every other rank: one grouped
// NCCL send / receive (a single fused kernel; capturable in the CG graph)
void MlPrec::exchange_slices(cudaStream_t st, double *buf, const std::vector<int64_t> &begin, int width)
{
    const size_t mine = (size_t)(width * (begin[(size_t)rank + 1] - begin[(size_t)rank]));
    NK(nc->GroupStart());
    for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        const size_t theirs = (size_t)(width * (begin[(size_t)q + 1] - begin[(size_t)q]));
        if (mine) NK(nc->Send(buf + width * begin[(size_t)rank], mine, nccl::kFloat64, q, comm, st));
        if (theirs) NK(nc->Recv(buf + width * begin[(size_t)q], theirs, nccl::kFloat64, q, comm, st));
    }
    NK(nc->GroupEnd());
}

void MlPrec::apply(cudaStream_t st, const double *rt, double *zt, double *part, int grid, const PcgScalars *sc, const PeerCtx *pc)
{
    const size_t L = lv.size();
    launch_ml_restrict0(st, agg_begin[(size_t)rank + 1] - agg_begin[(size_t)rank], group0, agg_ptr0.p, agg_rows0.p, t0_apply(), half16, rt,
                        lv[0]->r.p + 6 * agg_begin[(size_t)rank]);
    // several ranks: level 1 stays DISTRIBUTED in the apply -- every rank restricts its own level-1 rows to level 2 (partial
    // sums, all-reduced: 48 B per level-2 row) and prolongs back onto its own level-1 rows only, which is all the fine
    // prolongation needs; levels >= 2 are replicated.  With a single coarse level the dense solve needs all of r1.
    const bool split1 = world > 1 && L >= 2;
    if (world > 1 && !split1) exchange_slices(st, lv[0]->r.p, agg_begin, 6);
    for (size_t m = 0; m + 1 < L; ++m) {
        const int64_t lo = (m == 0 && split1) ? agg_begin[(size_t)rank] : 0, hi = (m == 0 && split1) ? agg_begin[(size_t)rank + 1] : lv[m]->n;
        if (lv[m]->smoothed)
            launch_ml_restrict(st, lv[m + 1]->n, lv[m]->nnzp, lv[m]->rrow.p, lv[m]->rblk.p, lv[m]->rfine.p, lv[m]->p_apply(half16), half16, lv[m]->r.p, lv[m + 1]->r.p, lo, hi);
        else
            launch_ml_restrict_tent(st, lv[m + 1]->n, lv[m]->rrow.p, lv[m]->rfine.p, lv[m]->off.p, lv[m]->r.p, lv[m + 1]->r.p, lo, hi);
        if (m == 0 && split1) NK(nc->AllReduce(lv[1]->r.p, lv[1]->r.p, (size_t)(6 * lv[1]->n), nccl::kFloat64, nccl::kSum, comm, st));
    }
    launch_ml_dense_matvec(st, (int)n_dense, dense32.p, lv[L - 1]->r.p, lv[L - 1]->z.p);
    for (size_t m = L - 1; m-- > 0;) {
        const int64_t row0 = (m == 0 && split1) ? agg_begin[(size_t)rank] : 0;
        const int64_t cnt = (m == 0 && split1) ? agg_begin[(size_t)rank + 1] - row0 : lv[m]->n;
        if (lv[m]->smoothed)
            launch_ml_smooth_prolong(st, row0, cnt, lv[m]->n, lv[m]->nnzp, lv[m]->dinv32.p, lv[m]->prow.p, lv[m]->pcol.p, lv[m]->p_apply(half16), half16, lv[m]->r.p,
                                     lv[m + 1]->z.p, lv[m]->z.p);
        else
            launch_ml_smooth_prolong_tent(st, row0, cnt, lv[m]->dinv32.p, lv[m]->agg.p, lv[m]->off.p, lv[m]->r.p, lv[m + 1]->z.p,
                                          lv[m]->z.p);
    }
    launch_ml_finish0(st, grid, nb, agg0.p, t0_apply(), half16, rt, lv[0]->z.p, zt, part, sc, pc, par.theta);
}

}  // namespace mbfea
