// The level loop of the additive multilevel preconditioner, written once against a "launcher" so that the GPU driver
// (api.cu, launch_ml_* on a stream) and the CPU emulation test (tests/multilevel_emul.cpp) run the SAME sequencing.
//
//   z~ = r~ + L^T B_0 [ sum over coarse levels ] B_0^T L r~,   level k+1 = aggregates of level k,
//   intermediate levels smoothed with D^-1 = S^T S of their Galerkin matrix, the last one solved with its dense inverse.
#pragma once
#include <cstddef>
#include <cstdint>

namespace mbfea {

struct MlLevelView {            // coarse level k (0-based): maps the rows of the level below it onto its own rows
    int64_t n_fine, n_coarse;   // rows below / rows of this level
    const int32_t *agg, *agg_ptr, *agg_rows;
    const double *off;
    const double *sinv;         // S of this level's diagonal blocks (unused on the last level)
    double *r, *z;              // 6 * n_coarse each
};

// Launcher concept: restrict(n_agg, agg_ptr, agg_rows, off, lfac_or_null, r, rc), smooth(n, sinv, r, z),
// dense_matvec(n, Ainv, r, z), prolong(n, agg, off, zc, z), finish(grid, nb, agg, off, lfac, rt, z1, zt, part)
template <class Launcher>
void ml_apply_levels(Launcher &&go, const MlLevelView *lv, size_t n_levels, const double *Ainv_last, int64_t nb,
                     const double *lfac, const double *rt, double *zt, double *part, int grid)
{
    const size_t L = n_levels;
    go.restrict(lv[0].n_coarse, lv[0].agg_ptr, lv[0].agg_rows, lv[0].off, lfac, rt, lv[0].r);
    for (size_t l = 1; l < L; ++l)
        go.restrict(lv[l].n_coarse, lv[l].agg_ptr, lv[l].agg_rows, lv[l].off, (const double *)nullptr, lv[l - 1].r, lv[l].r);
    go.dense_matvec(6 * lv[L - 1].n_coarse, Ainv_last, lv[L - 1].r, lv[L - 1].z);
    for (size_t l = L - 1; l >= 1; --l) {   // rows of level l-1 are the members of level l's aggregates
        go.smooth(lv[l - 1].n_coarse, lv[l - 1].sinv, lv[l - 1].r, lv[l - 1].z);
        go.prolong(lv[l - 1].n_coarse, lv[l].agg, lv[l].off, lv[l].z, lv[l - 1].z);
    }
    go.finish(grid, nb, lv[0].agg, lv[0].off, lfac, rt, lv[0].z, zt, part);
}

}  // namespace mbfea
